// Stand-in for the translation unit the reference's build system generates from
// util/build_version.cc.in (version strings + empty plugin registry).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#include <memory>
#include <string>
#include <unordered_map>

#include "rocksdb/utilities/object_registry.h"
#include "rocksdb/version.h"

namespace ROCKSDB_NAMESPACE {
std::unordered_map<std::string, RegistrarFunc> ObjectRegistry::builtins_ = {};

const std::unordered_map<std::string, std::string>& GetRocksBuildProperties() {
  static const std::unordered_map<std::string, std::string> props = {
      {"rocksdb_build_git_sha", "oracle-ref"}, {"rocksdb_build_date", "1970-01-01"}};
  return props;
}
std::string GetRocksVersionAsString(bool with_patch) {
  std::string v = std::to_string(ROCKSDB_MAJOR) + "." + std::to_string(ROCKSDB_MINOR);
  return with_patch ? v + "." + std::to_string(ROCKSDB_PATCH) : v;
}
std::string GetRocksBuildInfoAsString(const std::string& program, bool) {
  return program + " (RocksDB) " + GetRocksVersionAsString(true);
}
}  // namespace ROCKSDB_NAMESPACE
