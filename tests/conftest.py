import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "nextpolish1")
REF_SAMTOOLS = os.path.join(ROOT, "oracle", "_ref", "samtools")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def E():
    """nextpolish_b200.engine with the product library loaded (built by __graft_entry__.build())."""
    from nextpolish_b200 import binding, engine
    if not os.path.exists(binding.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    engine.lib()
    return engine


@pytest.fixture(scope="session")
def oracle():
    """The C restatement (oracle/libnp_oracle.so) — the checker, never the product."""
    path = os.path.join(ROOT, "oracle", "libnp_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    O = C.CDLL(path)
    O.np_oracle_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    O.np_oracle_run.restype = C.c_int
    return O


def _emu_build(name, defines=()):
    d = os.path.join(ROOT, "tests", "_emu")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, name)
    srcs = [os.path.join(ROOT, "tests", "emu", "emu_engine.cpp"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "engine_impl.h"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "engine_task2.h"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "engine_v2.h"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "diff_pass.h"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "column_pass.h"),
            os.path.join(ROOT, "nextpolish_b200", "csrc", "device_logic.h")]
    extra = [os.path.join(ROOT, "tests", "emu", "emu_bgzf.cpp"), os.path.join(ROOT, "nextpolish_b200", "csrc", "hostio.cpp")]
    srcs += extra + [os.path.join(ROOT, "nextpolish_b200", "csrc", "bgzf_inflate.h"), os.path.join(ROOT, "nextpolish_b200", "csrc", "hostio.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + list(defines) + ["-o", so, srcs[0]] + extra + ["-lz", "-lpthread"])
    L = C.CDLL(so)
    L.np_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.np_emu_run.restype = C.c_int
    L.np_emu_bgzf_inflate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.np_emu_bgzf_inflate.restype = C.c_int
    return L


@pytest.fixture(scope="session")
def emu():
    """Kernel bodies compiled for the host (tests/emu/emu_engine.cpp) — test build only."""
    return _emu_build("libnp_emu.so")


@pytest.fixture(scope="session")
def emu_general_slices():
    """The same test build with the column pass's 64-bit fast path limited to 9 columns per thread slice, so that
    ordinary inputs reach the general (many insertion sub-columns) loops of column_pass.h; and with task 2's region lists
    always merged by the literal per-contig walk (the fallback of the flat merge, engine_task2.h)."""
    return _emu_build("libnp_emu_slow.so", ("-DNP_SLICE_FAST_MAX=9", "-DNP_FORCE_REGION_WALK=1"))


def run_checker(fn, shard, task, cfg, extra=()):
    """Run oracle / emu entry `fn` over a shard -> {name: bytes}."""
    cap = int(shard.total_bases * 2) + 4096
    out = np.zeros(cap, np.uint8)
    off = np.zeros(shard.n_contigs + 1, np.int64)
    rc = fn(C.addressof(shard.view), task, C.cast(cfg, C.c_void_p), out.ctypes.data, cap, off.ctypes.data, *extra)
    assert rc == 0, "checker returned %d" % rc
    raw = out.tobytes()
    return {nm: raw[off[i]:off[i + 1]] for i, nm in enumerate(shard.names)}


def read_fasta(path):
    d, name = {}, None
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                name = line[1:].split()[0]
                d[name] = []
            elif name is not None:
                d[name].append(line.strip())
    return {k: "".join(v).encode() for k, v in d.items()}


def md5(b):
    return hashlib.md5(b).hexdigest()


@pytest.fixture(scope="session")
def synth_files(tmp_path_factory, E):
    """Writes the seeded synthetic FASTA+BAM of tests/synth_cases.py on demand; returns getter."""
    from tests.synth_cases import CASES
    base = tmp_path_factory.mktemp("synth")
    cache = {}

    def get(name):
        if name not in cache:
            fa, bam = str(base / (name + ".fa")), str(base / (name + ".bam"))
            p = E.synth_params(**CASES[name])
            assert E.lib().np_synth_write(p, fa.encode(), bam.encode()) == 0
            cache[name] = (fa, bam)
        return cache[name]
    return get
