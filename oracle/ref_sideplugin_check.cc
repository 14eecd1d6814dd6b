// oracle/ref_sideplugin_check.cc — test driver (test infrastructure): the reference's SidePlugin repository (sideplugin/rockside,
// compiled from the sources where they lie) imports a JSON document that names the B200 plugin classes, creates the objects through the
// factories the plugin registers with ROCKSDB_FACTORY_REG (plugin/*.cc, -DB200C_WITH_SIDEPLUGIN) and prints what it got.
//   usage: ref_sideplugin_check <config.json>
#include <cstdio>
#include <memory>
#include <string>

#include "db/compaction/compaction_executor.h"
#include "rocksdb/table.h"
#include "topling/side_plugin_repo.h"

using namespace ROCKSDB_NAMESPACE;

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <config.json>\n", argv[0]);
    return 2;
  }
  SidePluginRepo repo;
  Status s = repo.ImportJsonFile(argv[1]);
  if (!s.ok()) {
    fprintf(stderr, "import: %s\n", s.ToString().c_str());
    return 1;
  }
  std::shared_ptr<CompactionExecutorFactory> ex;
  std::shared_ptr<TableFactory> tf;
  const bool has_ex = repo.Get("b200", &ex), has_tf = repo.Get("b200_bbt", &tf);
  printf("{\"executor\": \"%s\", \"executor_allow_fallback\": %d, \"job_url\": \"%s\", \"table_factory\": \"%s\", \"delete_range\": %d}\n",
         has_ex && ex ? ex->Name() : "", has_ex && ex ? (int)ex->AllowFallbackToLocal() : -1,
         has_ex && ex ? ex->JobUrl("db", 7, 1).c_str() : "", has_tf && tf ? tf->Name() : "", has_tf && tf ? (int)tf->IsDeleteRangeSupported() : -1);
  return has_ex && has_tf ? 0 : 1;
}
