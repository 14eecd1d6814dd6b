"""Extracts the SingleDelete known-answer tests of the reference's CompactionJob test (db/compaction/compaction_job_test.cc:
SimpleSingleDelete, SingleDeleteSnapshots, EarliestWriteConflictSnapshot, SingleDeleteZeroSeq, MultiSingleDelete) into
tests/golden/compaction_job_kat.json: the mock input files with their levels, the expected output entries, the snapshot list and the
earliest write-conflict snapshot handed to RunCompaction.  Run where /root/reference exists; the JSON is what travels."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(REF, "db", "compaction", "compaction_job_test.cc")).read()
TYPES = {"kTypeValue": 1, "kTypeDeletion": 0, "kTypeSingleDeletion": 7}
cases = []
for name in ["SimpleSingleDelete", "SingleDeleteSnapshots", "EarliestWriteConflictSnapshot", "SingleDeleteZeroSeq", "MultiSingleDelete"]:
    start = src.index("TEST_F(CompactionJobTest, %s)" % name)
    body = src[start:src.index("\n}\n", start)]
    line = src[:start].count("\n") + 1
    files = {}
    for m in re.finditer(r"auto (\w+) =\s*mock::MakeMockFile\(\{(.*?)\}\);", body, re.S):
        ents = re.findall(r'\{KeyStr\("([^"]*)",\s*(\d+)U,\s*(kType\w+)\),\s*"([^"]*)"\}', m.group(2))
        assert len(ents) == m.group(2).count("KeyStr("), (name, m.group(1))
        files[m.group(1)] = [[k, int(s), TYPES[t], v] for k, s, t, v in ents]
    added = [(f, int(lvl) if lvl else 0) for f, lvl in re.findall(r"AddMockFile\((\w+)(?:,\s*(\d+))?\);", body)]
    rc = re.search(r"RunCompaction\(\{files\}, \{input_level\}, \{expected_results\}(?:,\s*\{([^}]*)\}(?:,\s*(\d+)U)?)?\);", body, re.S)
    assert rc, name
    snaps = [int(x.rstrip("U")) for x in rc.group(1).split(",") if x.strip()] if rc.group(1) else []
    cases.append({"name": name, "line": line,
                  "inputs": [{"level": lvl, "entries": files[f]} for f, lvl in added if lvl == 0],
                  "deeper_levels": [{"level": lvl, "entries": files[f]} for f, lvl in added if lvl > 0],
                  "expected": files["expected_results"], "snapshots": snaps,
                  "earliest_write_conflict_snapshot": int(rc.group(2)) if rc.group(2) else None})
here = os.path.dirname(os.path.abspath(__file__))
json.dump({"source": "db/compaction/compaction_job_test.cc (CompactionJobTest, output level 1, inputs = the level-0 mock files)",
           "cases": cases}, open(os.path.join(here, "compaction_job_kat.json"), "w"), indent=1)
for c in cases:
    print(c["name"], c["line"], [len(f["entries"]) for f in c["inputs"]], [len(f["entries"]) for f in c["deeper_levels"]], len(c["expected"]),
          c["snapshots"], c["earliest_write_conflict_snapshot"])
