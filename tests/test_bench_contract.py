"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs without a GPU, prints exactly one
JSON line on stdout with the keys the driver reads, and ranks other than 0 exit quietly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-log2", "6"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_json_line():
    lines = [l for l in run().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "jive_compressions_per_s" and d["unit"] == "compressions/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_reference_arm_does_not_map_the_product_library(tmp_path):
    """The reference arm must not load libanemoi_b200.so (nor import the package that dlopens it): only the oracle."""
    script = tmp_path / "probe.py"
    script.write_text(
        "import sys, types\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "args = types.SimpleNamespace(cpu_sample_log2=5, warmup=0, steps=1, gpus=1)\n"
        "bench.run_reference(args, 0)\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libanemoi_oracle' in maps\n"
        "assert 'libanemoi_b200' not in maps, 'reference arm mapped the product library'\n"
        "assert 'anemoi_rust_b200' not in sys.modules\n"
        "print('clean')\n" % ROOT)
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "clean" in out.stdout


def test_traffic_hash_agrees_with_the_profiling_tool():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, ROOT)
    import bench
    import ncu_summary

    assert bench.kernel_source_hash() == ncu_summary.kernel_source_hash()
