"""The worker's file conventions in the native library / CLI: block-file selection, resume scan of an
existing output part and the ">name_np<task> <len>" record header, against an independent statement of the rules of the
reference's lib/nextpolish1.py:148-179,226-229 (written out below) and against the Python worker mirror.  Host only."""
import ctypes as C
import os
import random
import subprocess

import pytest


# ---- the reference's rules, stated independently -------------------------------------------------------------------
def ref_scan_output(path):
    """read_polished_seqs (nextpolish1.py:163-179): finished names and the offset of the last record."""
    polished, last, offset, cur = set(), "", 0, 0
    with open(path, newline="") as f:
        for line in f:
            if line.startswith(">"):
                offset += cur
                cur = len(line)
                last = line.split()[0].split("_np")[0][1:]
                polished.add(last)
            else:
                cur += len(line)
    if last:
        polished.remove(last)
    return polished, offset


def ref_block_names(path, index, polished):
    """read_unpolished_seqs (nextpolish1.py:148-161), as an ordered list without duplicates."""
    out = []
    with open(path) as f:
        for line in f:
            if index != "all":
                p = line.strip().split()
                if p and len(p) > 1 and p[0].split("_np")[0] not in polished and p[1] == index and p[0] not in out:
                    out.append(p[0])
            elif line.startswith(">"):
                n = line.strip().split()[0][1:]
                if n.split("_np")[0] not in polished and n not in out:
                    out.append(n)
    return out


def ref_record_name(name, task):
    return name + (str(task) if name.split("_")[-1].startswith("np") else "_np" + str(task))     # nextpolish1.py:227


def plan(L, genome, block, index, out):
    enc = lambda s: s.encode() if s is not None else None
    p = L.np_part_plan_create(enc(genome), enc(block), enc(index), enc(out))
    assert p, L.np_last_error()
    names = [L.np_part_plan_name(p, i).decode() for i in range(L.np_part_plan_count(p))]
    r = (names, L.np_part_plan_finished(p), L.np_part_plan_resume_offset(p))
    L.np_part_plan_destroy(p)
    return r


@pytest.fixture()
def L(E):
    lib = E.lib()
    lib.np_last_error.restype = C.c_char_p
    return lib


def make_inputs(tmp_path, rng, n=12):
    names = ["ctg%03d" % i if i % 3 else "scaf_%d_np12" % i for i in range(n)]
    seqs = {nm: "".join(rng.choice("ACGTacgtN") for _ in range(rng.randrange(1, 400))) for nm in names}
    fa = str(tmp_path / "g.fa")
    with open(fa, "w") as f:
        for nm in names:
            f.write(">%s some description\n" % nm)
            s = seqs[nm]
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    blc = str(tmp_path / "g.blc")
    with open(blc, "w") as f:
        for i, nm in enumerate(names):
            f.write("%s\t%d\n" % (nm, i % 3))
        f.write("\n")                                       # a blank line is tolerated
    return names, seqs, fa, blc


def test_record_name_rule(L):
    buf = C.create_string_buffer(256)
    for nm in ("ctg1", "ctg1_np1", "ctg_np12", "a_b_c", "np", "x_npfoo", "x_np", "_", "ctg_np1_x"):
        for task in (1, 2, 4):
            n = L.np_part_record_name(nm.encode(), task, buf, 256)
            assert n == len(ref_record_name(nm, task)) and buf.value.decode() == ref_record_name(nm, task)
    assert L.np_part_record_name(b"abcdef", 1, buf, 5) == -1


def test_plan_without_output(L, tmp_path):
    rng = random.Random(1)
    names, _, fa, blc = make_inputs(tmp_path, rng)
    assert plan(L, fa, None, "all", None) == (names, 0, 0)
    assert plan(L, fa, blc, "all", "stdout") == (names, 0, 0)
    assert plan(L, fa, "", "1", None) == (names, 0, 0)                   # -i without -b: every contig (nextpolish1.py:212-214)
    for idx in ("0", "1", "2", "7"):
        assert plan(L, fa, blc, idx, str(tmp_path / "absent.fa")) == (ref_block_names(blc, idx, set()), 0, 0)
    p = L.np_part_plan_create(fa.encode(), str(tmp_path / "nope.blc").encode(), b"0", None)
    assert not p and b"cannot open" in L.np_last_error()


@pytest.mark.parametrize("seed", range(6))
def test_resume_scan_and_rewrite(L, tmp_path, seed):
    """Write some records through np_part_write, damage the tail in different ways, plan again: finished contigs are
    dropped, the last record is re-done, the file is cut where it starts; completing the part gives the full set."""
    rng = random.Random(seed)
    names, seqs, fa, blc = make_inputs(tmp_path, rng)
    out = str(tmp_path / "part.fasta")
    task = rng.choice([1, 2, 4])
    idx = str(rng.randrange(3))
    todo, nd, off = plan(L, fa, blc, idx, out)
    assert (todo, nd, off) == (ref_block_names(blc, idx, set()), 0, 0)
    k = rng.randrange(1, len(todo))
    f = L.np_part_open(out.encode(), 0)
    for nm in todo[:k]:
        assert L.np_part_write(f, nm.encode(), task, seqs[nm].encode(), len(seqs[nm]), 0) == 0
    assert L.np_part_close(f) == 0
    data = open(out).read()
    assert data == "".join(">%s %d\n%s\n" % (ref_record_name(nm, task), len(seqs[nm]), seqs[nm]) for nm in todo[:k])
    damage = seed % 3
    if damage == 1:                                         # killed in the middle of the last sequence
        data = data[:len(data) - 1 - rng.randrange(0, min(5, len(seqs[todo[k - 1]])))]
    elif damage == 2:                                       # killed right after a header
        data += ">%s %d\n" % (ref_record_name(todo[k], task), len(seqs[todo[k]]))
    open(out, "w").write(data)
    want_done, want_off = ref_scan_output(out)
    todo2, nd2, off2 = plan(L, fa, blc, idx, out)
    assert (todo2, nd2, off2) == (ref_block_names(blc, idx, want_done), len(want_done), want_off)
    assert todo2[0] == (todo[k] if damage == 2 else todo[k - 1])
    from nextpolish_b200 import nextpolish1 as mirror         # the Python worker mirror states the same rules
    pol = set()
    assert mirror.scan_output(out, pol) == off2 and mirror.read_block(blc, idx, pol) == todo2
    f = L.np_part_open(out.encode(), off2)
    for nm in todo2:
        assert L.np_part_write(f, nm.encode(), task, seqs[nm].encode(), len(seqs[nm]), 1 if seed == 3 else 0) == 0
    assert L.np_part_close(f) == 0
    want = "".join(">%s %d\n%s\n" % (ref_record_name(nm, task), len(seqs[nm]), seqs[nm].upper() if seed == 3 and nm in todo2 else seqs[nm])
                   for nm in todo)
    assert open(out).read() == want
    assert plan(L, fa, blc, idx, out)[0] == [todo[-1]]       # a complete part: only its last record is re-done


def test_long_lines_and_crlf(L, tmp_path):
    fa = str(tmp_path / "g.fa")
    big = "ACGT" * 100000                                   # one 400 kb line (longer than the reader's buffer)
    open(fa, "w").write(">a x\r\n%s\r\n>b\n%s\n" % (big, big))
    assert plan(L, fa, None, "all", None)[0] == ["a", "b"]
    out = str(tmp_path / "o.fa")
    open(out, "w", newline="").write(">a_np1 400000\n%s\n>b_np1 400000\n%s" % (big, big[:7]))
    names, nd, off = plan(L, fa, None, "all", out)
    assert (names, nd, off) == (["b"], 1, len(">a_np1 400000\n") + len(big) + 1)
    assert ref_scan_output(out) == ({"a"}, off)


def cli(E):
    return os.path.join(os.path.dirname(E.binding.LIB_PATH), "nextpolish1")


def test_cli_worker_grammar_without_gpu_work(E, tmp_path):
    """What the worker grammar does before any GPU work: argument errors, the refused tasks, and a job whose contigs are
    all finished already (nothing to polish: exit 0, the part is left as it is apart from its re-done last record)."""
    rng = random.Random(9)
    names, seqs, fa, blc = make_inputs(tmp_path, rng)
    exe = cli(E)
    r = subprocess.run([exe, "-g", fa, "-t", "5", "-s", "x.bam"], capture_output=True, text=True)
    assert r.returncode == 1 and "outside this engine" in r.stderr
    r = subprocess.run([exe, "-g", fa, "-t", "3", "-s", "x.bam", "-l", "y.bam"], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([exe, "-g", fa, "-s", "x.bam"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([exe, "-g", fa, "-t", "1", "-s", "x.bam", "-no_such_flag", "3"], capture_output=True, text=True)
    assert r.returncode == 2 and "unrecognized" in r.stderr
    # block 7 is empty: nothing to polish, no device touched, empty part created
    out = str(tmp_path / "part007.fasta")
    r = subprocess.run([exe, "-g", fa, "-t", "1", "-s", str(tmp_path / "absent.bam"), "-b", blc, "-i", "7", "-o", out, "-p", "4",
                        "-max_count_kmer", "30", "-max_variant_count_lgs", "150k", "-u"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == "" and "total time" in r.stderr


def test_cli_worker_fails_loudly_without_gpu(E, tmp_path):
    """A job with contigs to polish needs the device: on a box without one the worker exits 1 with the engine's message
    (no CPU path).  On a GPU box this test is skipped (the GPU suite covers the polished output)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tests.conftest import GOLDEN
    fa = os.path.join(GOLDEN, "td30.step1.fa")
    bam = os.path.join(GOLDEN, "td30.step1.bam")
    out = str(tmp_path / "p.fasta")
    r = subprocess.run([cli(E), "-g", fa, "-t", "1", "-s", bam, "-o", out], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr, r.stderr
