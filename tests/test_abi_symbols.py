"""CPU: libagym.so loads and exports every symbol include/agym.h declares (no compute calls without a GPU),
the ctypes mirror matches the header, and the product path fails loudly when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "agym.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+char\*|int64_t|uint64_t|size_t|int)\s+(agym_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_documented_entry_points():
    fns = declared_functions()
    for must in ("agym_create", "agym_destroy", "agym_simulate_rounds", "agym_replay_rounds", "agym_update_allocators",
                 "agym_k4_resolve", "agym_last_error", "agym_clear_iteration"):
        assert must in fns
    assert len(fns) >= 23


def test_library_exports_every_declared_symbol():
    import auction_gym_b200 as ag

    lib = ag._lib.load()  # raises if the .so is missing: build() must have run
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/agym.h but not exported by libagym.so"
        assert name in ag._lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(ag._lib.SIGNATURES) == set(declared_functions())
    assert lib.agym_abi_version() == ag._lib.ABI_VERSION == int(re.search(r"#define AGYM_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_ctypes_structs_match_header_layout():
    import auction_gym_b200 as ag

    L = ag._lib
    assert C.sizeof(L.Shape) == 10 * 4 + 8
    assert C.sizeof(L.RoundLog) == 15 * C.sizeof(C.c_void_p)
    assert C.sizeof(L.ReplayInputs) == 7 * C.sizeof(C.c_void_p) + 8
    hdr = open(HEADER).read()
    enum = re.search(r"enum agym_metric \{(.*?)\}", hdr, flags=re.S).group(1)
    names = [n.strip().split("=")[0].strip() for n in enum.replace("\n", " ").split(",") if n.strip()]
    assert names.index("AGYM_NUM_METRICS") == L.NUM_METRICS
    for n, v in (("AGYM_M_NET", L.M_NET), ("AGYM_M_BIAS", L.M_BIAS), ("AGYM_M_GAMMA", L.M_GAMMA), ("AGYM_M_NWON", L.M_NWON)):
        assert names.index(n) == v
    assert f"#define AGYM_BIDDER_D {L.BIDDER_D}" in hdr and f"#define AGYM_BIDDER_W {L.BIDDER_W}" in hdr


def test_no_cpu_fallback():
    import torch

    import auction_gym_b200 as ag

    if torch.cuda.is_available():
        pytest.skip("this check is for a box without a GPU")
    with pytest.raises(ag.AgymError, match="no CPU fallback"):
        ag.Engine(R=1, A=2, I=2, D=5, Do=4, P=2, mechanism=0, E=None, V=None, n_items=[2, 2], alloc_kind=[0, 0], bidder_kind=[0, 0])
    # and at the ABI itself
    lib = ag._lib.load()
    shape = ag._lib.Shape(1, 2, 2, 5, 4, 2, 0, 0, 0, 0, 1.0)
    h = C.c_void_p()
    assert lib.agym_create(C.byref(shape), 0, C.byref(h)) == -2  # AGYM_ERR_CUDA
    assert b"no CUDA device" in lib.agym_last_error(None)


def test_product_path_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import oracle/."""
    pkg = os.path.join(ROOT, "auction_gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle/"
                assert "/root/reference" not in text, f"{f} reads the reference tree"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/agym.h must be usable from C (no C++ / torch types): compile and run a C program against libagym.so."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(
        "#include <stdio.h>\n#include \"agym.h\"\n"
        "int main(void) {\n"
        "  agym_shape s = {1, 2, 2, 5, 4, 2, AGYM_SECOND_PRICE, AGYM_FP32, 0, 0, 1.0};\n"
        "  agym_handle* h = 0;\n"
        "  int bad = agym_create(0, 0, &h);                 /* null shape -> AGYM_ERR_INVALID */\n"
        "  s.P = 3; int bad2 = agym_create(&s, 0, &h);      /* P > A -> AGYM_ERR_INVALID */\n"
        "  printf(\"%d %d %d %s\\n\", agym_abi_version(), bad, bad2, agym_last_error(0));\n"
        "  return (agym_abi_version() == AGYM_ABI_VERSION && bad == AGYM_ERR_INVALID && bad2 == AGYM_ERR_INVALID) ? 0 : 1;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "auction_gym_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lagym", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("2 -1 -1 agym_create")  # AGYM_ABI_VERSION 2: agym_replay_inputs.num_slots, agym_shape.max_slots
