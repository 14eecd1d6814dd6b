"""SidePlugin registration of the two plugin classes (side_plugin_factory.h:290-293 ROCKSDB_FACTORY_REG): the reference's own plugin
repository -- sideplugin/rockside compiled from the sources where they lie (oracle/Makefile, target ref_sideplugin_check_mock) -- imports
a JSON document that names `B200Compact` and `B200BlockBasedTable` and hands back working objects; the JSON parameters reach them."""
import json
import os
import subprocess

import pytest

import helpers as H

BIN = os.path.join(H.ROOT, "oracle", "_ref", "ref_sideplugin_check_mock")
pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref not built (needs /root/reference)")


def _run(doc, tmp_path):
    cfg = tmp_path / "sideplugin.json"
    cfg.write_text(json.dumps(doc))
    r = subprocess.run([BIN, str(cfg)], capture_output=True, text=True)
    return r.returncode, (json.loads(r.stdout) if r.stdout.strip().startswith("{") else None), r.stderr


def test_repository_creates_the_plugin_objects_from_json(tmp_path):
    doc = {"CompactionExecutorFactory": {"b200": {"class": "B200Compact",
                                                   "params": {"device": 0, "allow_fallback_to_local": False, "scratch_dir": "/tmp/b200c"}}},
           "TableFactory": {"b200_bbt": {"class": "B200BlockBasedTable", "params": {"device": 0, "block_size": 8192, "format_version": 5}}}}
    rc, out, err = _run(doc, tmp_path)
    assert rc == 0, err
    assert out["executor"] == "B200Compact" and out["executor_allow_fallback"] == 0
    assert out["job_url"] == "b200c://cuda:0/db/job-7/att-1"
    assert out["table_factory"] == "B200BlockBasedTable" and out["delete_range"] == 1


def test_defaults_and_unknown_class(tmp_path):
    doc = {"CompactionExecutorFactory": {"b200": {"class": "B200Compact", "params": {}}},
           "TableFactory": {"b200_bbt": {"class": "B200BlockBasedTable", "params": {}}}}
    rc, out, err = _run(doc, tmp_path)
    assert rc == 0 and out["executor_allow_fallback"] == 1  # B200CompactOptions default
    doc["TableFactory"]["b200_bbt"]["class"] = "NoSuchTable"
    rc, out, err = _run(doc, tmp_path)
    assert rc != 0
