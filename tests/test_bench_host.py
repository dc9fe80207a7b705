"""CPU: host logic of bench.py that decides what the driver reads - the stale-capture refusal behind `roofline.traffic`
and the rank handling of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_traffic_record_is_refused_when_the_kernel_sources_changed(tmp_path, monkeypatch):
    now = bench.kernel_revision(bench.MSM_SOURCES)
    rec = {"k_accum_entries": {"kernel_revision": now, "dram_bytes_per_launch": 123.0, "algorithmic_bytes_per_launch": 45, "source": "test"}}
    f = tmp_path / "traffic.json"
    f.write_text(json.dumps(rec))
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(f))
    val, extra = bench.recorded_traffic("k_accum_entries", bench.MSM_SOURCES)
    assert val == 123.0 and extra["traffic_kernel_revision"] == now
    rec["k_accum_entries"]["kernel_revision"] = "0" * 16
    f.write_text(json.dumps(rec))
    val, extra = bench.recorded_traffic("k_accum_entries", bench.MSM_SOURCES)
    assert val is None and "stale capture refused" in extra["traffic_note"]
    val, extra = bench.recorded_traffic("k_ntt_pass", bench.NTT_SOURCES)      # no record for this kernel
    assert val is None and "no ncu capture" in extra["traffic_note"]


def test_committed_traffic_record_has_both_kernels():
    rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in ("k_accum_entries", "k_ntt_pass"):
        assert rec[k]["dram_bytes_per_launch"] > 0 and len(rec[k]["kernel_revision"]) == 16 and rec[k]["source"]


def test_kernel_revision_tracks_the_sources():
    a = bench.kernel_revision(bench.MSM_SOURCES)
    assert a == bench.kernel_revision(bench.MSM_SOURCES) and a != bench.kernel_revision(bench.NTT_SOURCES)


def test_reference_arm_runs_on_rank_zero_only():
    """under torchrun the other ranks of `--impl reference` exit 0 without work and without output"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
