"""Extracts the known-answer vectors of the reference's own Hash64 test (util/hash_test.cc, TEST(HashTest, Hash64SmallValueSchema):
`EXPECT_EQ(Hash64("<bytes>", <len>, kSeed), uint64_t{<value>u});`, kSeed = 0 = GetSliceHash64, the hash the Bloom filter builder feeds
on) into tests/golden/hash64_kat.json.  Run where /root/reference exists; the JSON is what travels."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(REF, "util", "hash_test.cc"), encoding="latin-1").read()
body = src[src.index("TEST(HashTest, Hash64SmallValueSchema)"):]
body = body[:body.index("\n}\n")]
body = body.replace('"\n', '"').replace("\n", " ")
out = []
for m in re.finditer(r'EXPECT_EQ\(\s*Hash64\(((?:\s*"(?:[^"\\]|\\.)*")+)\s*,\s*(\d+)\s*,\s*kSeed\s*\)\s*,\s*uint64_t\{(\d+)u\}\s*\)', body):
    lit, n, val = m.group(1), int(m.group(2)), int(m.group(3))
    data = b""
    for piece in re.findall(r'"((?:[^"\\]|\\.)*)"', lit):
        data += re.sub(rb"\\x([0-9a-fA-F]{2})", lambda h: bytes([int(h.group(1), 16)]), piece.encode("latin-1"))
    assert len(data) == n, (lit, n, data)
    out.append({"hex": data.hex(), "hash64": val})
here = os.path.dirname(os.path.abspath(__file__))
json.dump({"source": "util/hash_test.cc TEST(HashTest, Hash64SmallValueSchema), seed 0", "vectors": out},
          open(os.path.join(here, "hash64_kat.json"), "w"), indent=1)
print(len(out), "vectors; lengths", sorted({len(bytes.fromhex(v["hex"])) for v in out}))
