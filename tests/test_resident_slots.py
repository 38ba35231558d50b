"""np_resident (resident_slots.cu): several engines polishing shards that already sit in HBM at the same time;
every job's bytes must equal what one engine delivers for the same shard and task."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_view(E, torch, sh, with_qual):
    a = sh.arrays()
    keep = {k: torch.from_numpy(v.copy()).cuda() for k, v in a.items() if k in ("ctg_seq", "rec_off", "rec", "qual_off", "qual")}
    v = E.ShardView()
    v.n_contigs, v.n_reads = sh.view.n_contigs, sh.view.n_reads
    v.ctg_off, v.ctg_read_off = sh.view.ctg_off, sh.view.ctg_read_off
    v.ctg_seq, v.rec_off, v.rec = keep["ctg_seq"].data_ptr(), keep["rec_off"].data_ptr(), keep["rec"].data_ptr()
    if with_qual:
        v.qual_off, v.qual = keep["qual_off"].data_ptr(), keep["qual"].data_ptr()
    return v, keep


@pytest.mark.parametrize("slots", [1, 3])
def test_resident_slots_equal_one_engine(E, slots):
    import torch
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    eng = E.Engine(0)
    shards = [E.Shard.synthetic(E.synth_params(seed=700 + k, n_contigs=2 + k, contig_len=120000, depth=25.0,
                                               lowercase_frac=0.002 * (k + 1)), 0, 2 + k, with_qual=2) for k in range(3)]
    views = [_device_view(E, torch, sh, True) for sh in shards]
    jobs = [(t, k) for k in range(len(shards)) for t in E.TASKS] * 3
    want = {}
    for t, k in set(jobs):
        got = eng.polish(shards[k], t, cfg)
        want[(t, k)] = b"".join(got[n] for n in shards[k].names)
    eng.close()
    cap = int(max(sh.total_bases for sh in shards) * 2) + 4096
    rp = E.ResidentSlots(0, slots)
    bufs = [torch.zeros(cap + 16, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    pending = []

    def collect():
        tk, t, k = pending.pop(0)
        n = rp.wait(tk)
        b = bufs[tk % slots]
        assert int(b[:8].cpu().view(torch.int64).item()) == n == len(want[(t, k)])
        assert b[16:16 + n].cpu().numpy().tobytes() == want[(t, k)], (t, k, slots)

    for i, (t, k) in enumerate(jobs):            # tickets are consecutive from 0: job i lands in slot i % slots
        while len(pending) >= slots:
            collect()
        tk = rp.submit(t, views[k][0], cfg, bufs[i % slots].data_ptr(), cap + 16)
        assert tk == i
        pending.append((tk, t, k))
    while pending:
        collect()
    assert rp.launch_count() > 0
    with pytest.raises(E.NativeError):
        rp.wait(10 ** 6)
    rp.close()


def test_files_pipeline_queue_matches_reference_md5(E, synth_files):
    """np_files: workers pull jobs from a queue; 2 x depth jobs may be outstanding, results come back per ticket and equal
    the committed md5s of the reference binary's output on the same files (needs <bam>.bai: the GPU loader)."""
    import json
    import os
    import subprocess
    from tests.conftest import GOLDEN, REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("no samtools to index the BAMs")
    cases = ["c30", "lower", "ragged"]
    want = json.load(open(os.path.join(GOLDEN, "synth_md5.json")))
    files = {}
    for c in cases:
        fa, bam = synth_files(c)
        if not os.path.exists(bam + ".bai"):
            subprocess.check_call([REF_SAMTOOLS, "index", bam])
        files[c] = (fa, bam)
    fp = E.FilePipeline(0, depth=2)
    assert fp.capacity == 4
    jobs = [(c, t) for c in cases for t in (1, 2)] * 2
    done = 0

    def collect():
        nonlocal done
        c, t = jobs[done]
        r = fp.wait_oldest(want_md5=True)
        assert r["task"] == t
        assert {"%s_%d" % (n, t): h for n, h in r["md5"].items()} == want[c][str(t)], (c, t)
        done += 1

    for c, t in jobs:
        while fp.in_flight() > fp.capacity - 1:
            collect()
        fp.submit(t, files[c][0], files[c][1], E.default_config(files[c][0], files[c][1]))
    with pytest.raises(E.NativeError):              # every record holds an uncollected job
        while True:
            fp.submit(1, files["c30"][0], files["c30"][1], E.default_config(files["c30"][0], files["c30"][1]))
            jobs.append(("c30", 1))
    while fp.in_flight():
        collect()
    fp.close()
