"""GPU inflate of BGZF: the warp decoder of nextpolish_b200/csrc/bgzf_inflate.h against zlib.
CPU: the decoder body through the one-lane test backend (tests/emu/emu_bgzf.cpp).  GPU: np_bgzf_inflate."""
import ctypes as C
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from tests.conftest import GOLDEN


def bgzf_block(payload, level=6, wbits=-15, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
    d = co.compress(payload) + co.flush()
    bsize = len(d) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + d +
            struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))


def synthetic_streams():
    rng = np.random.default_rng(7)
    text = (b"ACGTTTGACCA" * 3000 + bytes(rng.integers(65, 90, 20000, dtype=np.uint8)))[:60000]
    rnd = bytes(rng.integers(0, 256, 30000, dtype=np.uint8))
    out = {}
    out["levels"] = b"".join(bgzf_block(text, lv) for lv in (0, 1, 6, 9)) + bgzf_block(b"")
    out["fixed_huffman"] = b"".join(bgzf_block(text[i:i + 90], 6, strategy=zlib.Z_FIXED) for i in range(0, 3000, 90))
    out["incompressible"] = bgzf_block(rnd, 6) + bgzf_block(rnd[:1], 6) + bgzf_block(rnd[:2], 0)
    out["long_matches"] = bgzf_block(b"A" * 65000, 9) + bgzf_block(b"AB" * 30000, 6) + bgzf_block((b"xyz" * 7 + b"Q") * 2900, 1)
    out["huffman_only"] = bgzf_block(text, 6, strategy=zlib.Z_HUFFMAN_ONLY) + bgzf_block(text, 6, strategy=zlib.Z_RLE)
    return out


def zlib_inflate(comp):
    """Reference result: every BGZF block inflated by zlib (raw deflate payload between header and trailer)."""
    out, off = [], 0
    while off < len(comp):
        xlen = struct.unpack_from("<H", comp, off + 10)[0]
        bsize = struct.unpack_from("<H", comp, off + 16)[0] + 1        # BC subfield first (htslib / our writers)
        assert comp[off + 12:off + 14] == b"BC"
        out.append(zlib.decompress(comp[off + 12 + xlen:off + bsize - 8], -15))
        off += bsize
    return b"".join(out)


def inflate_with(fn, comp, *lead):
    buf = np.frombuffer(comp, dtype=np.uint8)
    n = C.c_int64(0)
    nb = C.c_int32(0)
    assert fn(*lead, buf.ctypes.data, len(comp), None, 0, C.byref(n), C.byref(nb)) == 0
    out = np.zeros(max(n.value, 1), np.uint8)
    rc = fn(*lead, buf.ctypes.data, len(comp), out.ctypes.data, n.value, C.byref(n), C.byref(nb))
    return rc, out[:n.value].tobytes(), nb.value


@pytest.mark.parametrize("name", sorted(synthetic_streams()))
def test_emulated_decoder_matches_zlib_on_synthetic_blocks(emu, name):
    comp = synthetic_streams()[name]
    rc, got, nb = inflate_with(emu.np_emu_bgzf_inflate, comp)
    assert rc == 0
    assert got == zlib_inflate(comp)


@pytest.mark.parametrize("f", ["td30.step1.bam", "td30.step2.bam", "td30.step1.bam.bai"])
def test_emulated_decoder_matches_zlib_on_golden_bam(emu, f):
    comp = open(os.path.join(GOLDEN, f), "rb").read()
    if f.endswith(".bai"):
        comp = bgzf_block(comp[:60000], 6)       # an index is not BGZF: wrap it (binary payload)
    rc, got, nb = inflate_with(emu.np_emu_bgzf_inflate, comp)
    assert rc == 0 and nb > 0
    assert got == zlib_inflate(comp)


@pytest.mark.parametrize("seed", range(4))
def test_emulated_decoder_fuzz(emu, seed):
    """Random payload kinds x sizes (0 .. 65280) x zlib levels / strategies, several blocks per stream."""
    import random
    rng = random.Random(500 + seed)
    nrng = np.random.default_rng(rng.randrange(1 << 30))

    def payload():
        kind = rng.choice(["rand", "text", "runs", "mixed", "alphabet", "zeros"])
        n = rng.choice([0, 1, 2, 3, 7, 63, 64, 65, 255, 256, 257, 258, 259, 1000, 4095, 4096, 5000, 20000, 65000, 65280])
        if kind == "rand":
            return bytes(nrng.integers(0, 256, n, dtype=np.uint8))
        if kind == "text":
            return (b"ACGTTTGACCAGGT" * 6000)[:n]
        if kind == "runs":
            return b"".join(bytes([rng.randrange(256)]) * rng.choice([1, 2, 3, 4, 5, 30, 258, 259, 600]) for _ in range(300))[:n]
        if kind == "mixed":
            return bytes(nrng.choice([65, 67, 71, 84, 10, 33], n).astype(np.uint8))
        if kind == "alphabet":
            return bytes(nrng.integers(0, rng.choice([2, 3, 5, 17, 100]), n, dtype=np.uint8))
        return b"\0" * n

    for _ in range(25):
        comp = b"".join(bgzf_block(payload(), rng.choice(range(10)), strategy=rng.choice(
            [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])) for _ in range(rng.choice([1, 2, 5])))
        rc, got, nb = inflate_with(emu.np_emu_bgzf_inflate, comp)
        assert rc == 0 and got == zlib_inflate(comp)


def test_emulated_decoder_rejects_corrupt_payload(emu):
    comp = bytearray(bgzf_block(b"hello hello hello hello" * 100, 6))
    comp[30] ^= 0x55
    rc, got, nb = inflate_with(emu.np_emu_bgzf_inflate, bytes(comp))
    assert rc != 0 or got != b"hello hello hello hello" * 100      # never a silent success with the right bytes


def _gpu_fn(E):
    L = E.lib()

    def fn(dev, comp, n, out, cap, nout, nb):
        return L.np_bgzf_inflate(dev, comp, n, out, cap, nout, nb, None)
    return fn


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(synthetic_streams()))
def test_gpu_inflate_matches_zlib_on_synthetic_blocks(E, name):
    comp = synthetic_streams()[name]
    rc, got, nb = inflate_with(_gpu_fn(E), comp, 0)
    assert rc == 0, E.last_error()
    assert got == zlib_inflate(comp)


@pytest.mark.gpu
def test_gpu_inflate_matches_zlib_on_bams(E, synth_files):
    files = [os.path.join(GOLDEN, "td30.step1.bam"), os.path.join(GOLDEN, "td30.step2.bam"), synth_files("c30")[1], synth_files("noisy")[1]]
    for f in files:
        comp = open(f, "rb").read()
        rc, got, nb = inflate_with(_gpu_fn(E), comp, 0)
        assert rc == 0, E.last_error()
        assert nb > 1 and got == zlib_inflate(comp), f


def _bam_record_voffsets(comp):
    """Independent walk (zlib + struct): virtual offset and reference id of every alignment record of a BAM."""
    blocks, off, uoff = [], 0, 0
    while off < len(comp):
        xlen = struct.unpack_from("<H", comp, off + 10)[0]
        bsize = struct.unpack_from("<H", comp, off + 16)[0] + 1
        data = zlib.decompress(comp[off + 12 + xlen:off + bsize - 8], -15)
        blocks.append((off, uoff, len(data)))
        uoff += len(data)
        off += bsize
    u = zlib_inflate(comp)
    l_text = struct.unpack_from("<i", u, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", u, p)[0]
    p += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", u, p)[0]
        p += 4 + l_name + 4
    import bisect
    ustarts = [b[1] for b in blocks]
    recs = []
    while p + 4 <= len(u):
        bs = struct.unpack_from("<i", u, p)[0]
        tid = struct.unpack_from("<i", u, p + 4)[0]
        k = bisect.bisect_right(ustarts, p) - 1
        while blocks[k][2] == 0 or p - blocks[k][1] >= blocks[k][2]:      # skip empty blocks / move to the owning block
            k += 1
        recs.append(((blocks[k][0] << 16) | (p - blocks[k][1]), tid))
        p += 4 + bs
    return recs


@pytest.mark.parametrize("f", ["td30.step1.bam", "td30.step2.bam"])
def test_index_anchors_are_record_starts(emu, f):
    """The premise of the device-side record walk (devload.cu): every virtual offset BamFile::bai_record_starts takes
    from the .bai (chunk begins + linear index) is the start of a record of that reference, and the first record of
    every reference is among them."""
    path = os.path.join(GOLDEN, f)
    recs = _bam_record_voffsets(open(path, "rb").read())
    emu.np_emu_bai_record_starts.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_int64]
    emu.np_emu_bai_record_starts.restype = C.c_int64
    by_tid = {}
    for v, tid in recs:
        by_tid.setdefault(tid, []).append(v)
    assert by_tid
    for tid, vs in by_tid.items():
        if tid < 0:
            continue
        out = (C.c_uint64 * 100000)()
        n = emu.np_emu_bai_record_starts(path.encode(), tid, out, 100000)
        assert n > 0
        anchors = [out[i] for i in range(n)]
        assert anchors == sorted(set(anchors))
        assert set(anchors) <= set(vs), "an index offset is not a record start of reference %d" % tid
        assert anchors[0] == vs[0]
