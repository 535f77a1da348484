"""CPU suite: the parts of bench.py that do not need a GPU -- the ncu captures bench.py reads `roofline.traffic` from are
committed and parse, the workload description names the mask-row mode, the command line parses."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_committed_ncu_captures_parse():
    b = load_bench()
    full, per_full = b.ncu_traffic(False)
    lazy, per_lazy = b.ncu_traffic(True)
    # the overlapped step with every row zeroed in full: ~87 MB per step, 3/4 of it row writes; with lazy clearing: ~24 MB
    assert 80e6 < full < 95e6 and per_full["dram_write_per_step"] > 60e6
    assert 15e6 < lazy < 35e6 and per_lazy["dram_write_per_step"] < 15e6
    assert lazy < full / 3


def test_workload_config_names_the_mask_row_mode():
    b = load_bench()
    cfg = b.workload_config(100_000_000, 4096)
    assert "workload" in cfg and "mask_rows" in cfg and "model" not in cfg


def test_command_line_parses():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "--gpus" in r.stdout + r.stderr
