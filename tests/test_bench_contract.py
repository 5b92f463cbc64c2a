"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with
the agreed keys, and the workload table matches BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "blocks/s" and j["higher_is_better"] is True
    assert j["value"] > 1e6 and j["steps"] == 1 and j["warmup"] >= 3
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "BC1 8192x8192" in j["config"]["workload"]
    # both arms print the SAME config dict for the same command line (same_config must read true)
    sys.path.insert(0, ROOT)
    import bench
    assert j["config"] == bench.config_dict(bench.HEADLINE, 1)
    assert "warm output buffer" in j["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workloads_cover_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    wl = bench.WORKLOADS
    assert wl["bc1"]["parts"][0][:3] == ("BC1", 8192, 8192)            # configs[1]
    assert wl["bc7"]["parts"][0][:3] == ("BPTC", 16384, 16384)         # configs[2]
    assert wl["bc6h"]["parts"][0][:3] == ("BPTC_FLOAT", 8192, 8192)    # configs[3]
    assert [p[0] for p in wl["mixed"]["parts"]] == ["ETC2", "EAC_RG11", "RGTC2"] and wl["mixed"]["total_blocks"] == 10 ** 8
    # the default run carries the whole metric: headline + per_config + strong_1e8 legs
    assert bench.HEADLINE == "bc1" and wl["bc1_1e8"]["total_blocks"] == 10 ** 8 and wl["bc1_1e8"]["strong"] and wl["mixed"]["strong"]
    assert sum(bench.part_share(wl["mixed"], i) for i in range(3)) == 10 ** 8
    assert [bench.part_share(wl["mixed"], i) for i in range(3)] == [33333334, 33333333, 33333333]
    for fmt, b in (("BC1", 72), ("BPTC", 80), ("BPTC_FLOAT", 144), ("RGTC2", 48), ("EAC_RG11", 80), ("ETC2", 72)):
        assert bench.ALGO_BYTES[fmt] == b
