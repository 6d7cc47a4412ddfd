"""The C-ABI library loads and exports exactly the symbols include/b200ssl.h declares (no GPU needed)."""
import ctypes
import os
import re

from cv_ssl_mis_b200 import _lib


def header_functions():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in b200ssl.h but not exported"
    assert _lib.load().b200_abi_version() == _lib.ABI_VERSION


def test_no_torch_types_in_the_abi():
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)      # declarations only
    assert "torch" not in src.lower() and "at::" not in src and "Tensor" not in src


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(_lib.__file__))
    pkg = os.path.join(root, "cv_ssl_mis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text or "import oracle" not in text and "from oracle" not in text, f


def test_reference_arm_prints_the_bench_contract():
    """`bench.py --impl reference` (the oracle port on the host cores) prints one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(_lib.__file__))
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "slices/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "MT-UNet" in line["metric"] and "workload" in line["config"]


def test_every_op_has_a_cpu_stand_in():
    """tests/fake_ops.py mirrors the whole ops surface, so every launch schedule can be exercised without a GPU."""
    from cv_ssl_mis_b200 import ops
    from tests import fake_ops
    skip = {"ConvDesc", "B200Error", "conv_desc", "desc_out_dims"}
    missing = [n for n in dir(ops) if not n.startswith("_") and callable(getattr(ops, n)) and n not in skip
               and not hasattr(fake_ops, n)]
    assert not missing, missing
