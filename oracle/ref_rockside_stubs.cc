// oracle/ref_rockside_stubs.cc — link stand-ins (test infrastructure) for the parts of sideplugin/rockside that the registration test
// (ref_sideplugin_check) never calls: the embedded web server, the YAML front end, the DB-opening helpers and the option printers.  The
// repository itself (side_plugin_repo.cc), the factory templates (side_plugin_tpl_inst.cc) and the table-factory plugins are the
// reference's own sources, compiled where they lie.  Every stand-in aborts if it is ever reached.
#include <cstdio>
#include <cstdlib>

#include "topling/side_plugin_internal.h"
#include "topling/side_plugin_repo.h"
#include "topling/web/json_civetweb.h"

namespace ROCKSDB_NAMESPACE {
[[noreturn]] static void Unreached(const char* what) {
  fprintf(stderr, "ref_rockside_stubs: %s is not part of this test build\n", what);
  abort();
}
JsonCivetServer::JsonCivetServer() : m_impl(nullptr) {}
JsonCivetServer::~JsonCivetServer() {}
void JsonCivetServer::Init(const json&, SidePluginRepo*) { Unreached("JsonCivetServer::Init"); }
void JsonCivetServer::Close() {}
std::string YamlToJson(std::string&) { Unreached("YamlToJson"); }
json JsonFromText(const std::string& text) { return json::parse(text); }
json DBOptionsToJson(const DBOptions&, const SidePluginRepo&) { return json::object(); }
json CFOptionsToJson(const ColumnFamilyOptions&, const SidePluginRepo&) { return json::object(); }
void DynaMemTableBackPatch(MemTableRepFactory*, const SidePluginRepo&) {}
void JS_ToplingDB_AddVersion(json&, bool) {}
void JS_TopTable_AddVersion(json&, bool) {}
void JS_CSPPMemTab_AddVersion(json&, bool) {}
void JS_CSPP_WBWI_AddVersion(json&, bool) {}
void JS_ToplingDcompact_AddVersion(json&, bool) {}

DB_MultiCF::DB_MultiCF() {}
DB_MultiCF::~DB_MultiCF() {}
DB_MultiCF_Impl::DB_MultiCF_Impl(const SidePluginRepo*, const std::string&, DB*, const std::vector<ColumnFamilyHandle*>&, int) {
  Unreached("DB_MultiCF_Impl");
}
DB_MultiCF_Impl::DB_MultiCF_Impl() { Unreached("DB_MultiCF_Impl"); }
DB_MultiCF_Impl::~DB_MultiCF_Impl() {}
ColumnFamilyHandle* DB_MultiCF_Impl::Get(const std::string&) const { Unreached("DB_MultiCF_Impl::Get"); }
Status DB_MultiCF_Impl::CreateColumnFamily(const std::string&, const std::string&, ColumnFamilyHandle**) { Unreached("CreateColumnFamily"); }
Status DB_MultiCF_Impl::DropColumnFamily(const std::string&, bool) { Unreached("DropColumnFamily"); }
Status DB_MultiCF_Impl::DropColumnFamily(ColumnFamilyHandle*, bool) { Unreached("DropColumnFamily"); }
std::vector<ColumnFamilyHandle*> DB_MultiCF_Impl::get_cf_handles_view() const { Unreached("get_cf_handles_view"); }
}  // namespace ROCKSDB_NAMESPACE
