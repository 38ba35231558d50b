"""Host-side readers of the loader (csrc/hostio.cpp) against independent parses — CPU only (test build tests/emu)."""
import ctypes as C
import os
import random

import numpy as np
import pytest


def _ref_parse(data: bytes):
    """FASTA rules of the loader: name = header up to the first white space; a sequence keeps the isgraph bytes (33..126)
    of its lines, case preserved; anything before the first header is skipped."""
    names, seqs, cur = [], [], None
    for line in data.split(b"\n"):
        if line.startswith(b">"):
            names.append(line[1:].split()[0].decode() if line[1:].split() else "")
            cur = bytearray()
            seqs.append(cur)
        elif cur is not None:
            cur.extend(b for b in line if 33 <= b <= 126)
    return names, [bytes(s) for s in seqs]


def _load(emu, path, cap=1 << 22):
    emu.np_emu_fasta_flat.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    emu.np_emu_fasta_flat.restype = C.c_int64
    seq = np.zeros(cap, np.uint8)
    off = np.zeros(4096, np.int64)
    names = C.create_string_buffer(1 << 16)
    n = C.c_int32(0)
    total = emu.np_emu_fasta_flat(str(path).encode(), seq.ctypes.data, cap, off.ctypes.data, 4095, names, len(names), C.byref(n))
    assert total >= 0, total
    nm = names.value.decode().split("\n")[:-1] if n.value else []
    raw = seq.tobytes()
    return nm, [raw[off[i]:off[i + 1]] for i in range(n.value)]


CASES = {
    "plain": b">a desc\nACGT\nacgt\n>b\nNNNN\n",
    "no_trailing_newline": b">a\nACGT\n>b\nGG",
    "empty_lines_and_spaces": b">a\n\nAC GT\n\n  \n>b\t x\nT T\n",
    "crlf": b">a\r\nACGT\r\nAC\r\n>b\r\nGG\r\n",
    "junk_before_header": b"garbage\nmore\n>a\nAC\n",
    "header_only": b">a\n>b\nAC\n>c\n",
    "empty": b"",
    "only_junk": b"no header here\n",
    "lowercase_and_iupac": b">x\nacgtnRYKMswbdhv\nACGT*-\n",
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_fasta_load_flat_edge_cases(emu, tmp_path, case):
    p = tmp_path / (case + ".fa")
    p.write_bytes(CASES[case])
    assert _load(emu, p) == _ref_parse(CASES[case])


@pytest.mark.parametrize("seed", range(4))
def test_fasta_load_flat_random(emu, tmp_path, seed):
    """Random contig counts, line widths and lengths (some longer than a page, some empty); the same file is read twice:
    the reader keeps its buffer between calls."""
    rng = random.Random(seed)
    out = bytearray()
    for i in range(rng.randint(1, 40)):
        n = rng.choice([0, 1, 59, 60, 61, 4095, 4096, 4097, rng.randint(0, 50000)])
        w = rng.choice([1, 60, 80, 1000000])
        s = bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
        out += b">ctg%d some description\n" % i
        for k in range(0, n, w):
            out += s[k:k + w] + b"\n"
    p = tmp_path / "r.fa"
    p.write_bytes(bytes(out))
    want = _ref_parse(bytes(out))
    assert _load(emu, p) == want
    q = tmp_path / "small.fa"
    q.write_bytes(b">s\nAC\n")
    assert _load(emu, q) == (["s"], [b"AC"])
    assert _load(emu, p) == want


def test_host_reader_checks_the_bgzf_crc(E, tmp_path):
    """A flipped bit in a block's CRC32 trailer (the payload still inflates to the right length) is an error, as in htslib."""
    import shutil
    from tests.conftest import GOLDEN
    src = os.path.join(GOLDEN, "td30.step1.bam")
    fa = os.path.join(GOLDEN, "td30.step1.fa")
    bad = str(tmp_path / "crc.bam")
    raw = bytearray(open(src, "rb").read())
    # second BGZF block: header 18 bytes (BC subfield holds the block size - 1), trailer = CRC32, ISIZE
    bsize0 = raw[16] | (raw[17] << 8)
    o = bsize0 + 1
    assert raw[o:o + 4] == b"\x1f\x8b\x08\x04"
    bsize1 = raw[o + 16] | (raw[o + 17] << 8)
    raw[o + bsize1 + 1 - 8] ^= 0x01
    open(bad, "wb").write(raw)
    shutil.copy(src + ".bai", bad + ".bai")
    with pytest.raises(E.NativeError):
        E.Shard.load(fa, bad, with_qual=False)
    E.Shard.load(fa, src, with_qual=False).close()
