#!/usr/bin/env python
"""bench.py — Mbp polished/s of the polishing hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1], "c2"): synthetic 5 Mb draft (5 contigs x 1 Mb) + 30x 150 bp PE short
reads per GPU; a "step" = one pass of every implemented task step (score_chain, kmer_count) over one such
shard.  Weak scaling: every rank polishes its own shard of that shape (contigs are independent units,
SURVEY.md 8e); after each task step the polished FASTA bytes are gathered to rank 0 with one NCCL collective.

value        device-resident: packed shards already in HBM when the timed region starts, polished through
             np_resident_submit / np_resident_wait — several engines (stream + scratch + host thread each) work on
             different shards at once; `resident.one_engine` is the same K steps on ONE engine and stream
e2e          FROM THE FILES the reference reads: draft FASTA + BGZF BAM (+ .bai) in host memory (page cache) ->
             np_files_submit / np_files_wait (the pipelined form of what the drop-in score_chain()/kmer_count()
             and the native CLI do): compressed bytes host -> device, BGZF inflate + record unpack + packing on
             the GPU, polishing kernels, polished bytes device -> host.  This is what the reference arm does on
             the CPU with the same files, so e2e / reference is an apples-to-apples ratio.  The pipeline is warmed
             with 4 x workers untimed steps, then exactly K steps are timed (`wall_ms_at_quarters`: drift inside).
e2e_packed   informational: the same step from pre-packed shards in pinned host memory (np_stream_*)
roofline     the pileup-scan kernel = the kernel that streams the sorted read blocks against the draft (k_diff,
             "pileup_diff"): algorithmic bytes / its CUDA-event time on the engine stream, against
             the measured HBM peak of MEASURED_PEAKS.json; `pair` adds the column kernel that consumes its output
             (k_col_pass, "pileup_scan"), `task1` gives the same ratio for ALL task-1 kernels together
cpu_baseline / --impl reference: the reference's own CPU implementation (oracle/_ref/nextpolish1 compiled from
             the reference sources), one process per contig like nextpolish1.py's Pool and as many concurrent
             copies of the step as the host has cores for (all host threads busy), else the oracle port.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n_contigs=5, contig_len=1000000, depth=30.0, read_len=150)
SEED0 = 20240917 + 2
# host threads per rank: every engine slot / pipeline job is driven by its own host thread, so both scale with the cores a rank gets
_CPUS_PER_RANK = max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
RES_SLOTS = int(os.environ.get("NP_BENCH_SLOTS", str(max(2, min(8, _CPUS_PER_RANK // 2)))))       # engines working concurrently on resident shards (np_resident)
FILES_DEPTH = int(os.environ.get("NP_BENCH_FILES_DEPTH", str(max(2, min(6, (3 * _CPUS_PER_RANK) // 8)))))   # jobs in flight in the from-files pipeline (host parse + upload of one job overlap the kernels of the others)
N_ROTATE = 3          # distinct resident shards rotated between steps (defeats L2 reuse across steps)
# the task-2 step runs on what the pipeline hands it: reads re-mapped to the task-1 output, i.e. a nearly clean draft
# (residual error 1e-5 / 2e-5) whose unsupported bases are lowercase.  lowercase_frac = 6.3e-4 is what the reference's
# task 1 leaves on this generator's 30x shards (measured on tests/synth_cases.py c30); real data can be denser
# (4.0e-3 on tests/golden/td30.step1.expected.fa): `task2_sensitivity` in the JSON line times that case too.
TASK2_DRAFT = dict(draft_snv=1e-5, draft_indel=2e-5, lowercase_frac=6.3e-4)
TASK2_DENSE = dict(draft_snv=1e-5, draft_indel=2e-5, lowercase_frac=4.0e-3)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "nextpolish1")
SAMTOOLS = os.path.join(ROOT, "oracle", "_ref", "samtools")
SIMULATE = os.path.join(ROOT, "nextpolish_b200", "lib", "np_simulate")
CMDS = {1: "scorechain", 2: "kmercount"}


def rank_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def synth_kwargs(task, seed, dense=False):
    extra = dict(lowercase_frac=0.0) if task == 1 else (TASK2_DENSE if dense else TASK2_DRAFT)
    d = dict(seed=seed, **WORKLOAD)
    d.update(extra)
    return d


def seed_for(rank, task, k=0):
    return SEED0 + 1000 * rank + k + 100 * task


def write_inputs(tmpdir, rank, tasks):
    """FASTA + BAM + .bai of the workload for `rank` (written by the standalone generator binary: the product
    library is not involved).  Returns {task: (fasta, bam)}."""
    files = {}
    for t in tasks:
        fa, bam = os.path.join(tmpdir, "c2.r%d.t%d.fa" % (rank, t)), os.path.join(tmpdir, "c2.r%d.t%d.bam" % (rank, t))
        kw = synth_kwargs(t, seed_for(rank, t))
        subprocess.check_call([SIMULATE, fa, bam] + ["%s=%r" % (k, v) for k, v in kw.items()])
        subprocess.check_call([SAMTOOLS, "index", bam])
        files[t] = (fa, bam)
    os.sync()       # the inputs sit in the page cache; their write-back must not run underneath the timed reads
    return files


def read_fasta_bytes(path):
    d, name = {}, None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                name = line[1:].split()[0].decode()
                d[name] = []
            elif name is not None:
                d[name].append(line.strip())
    return {k: b"".join(v) for k, v in d.items()}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML every 5 ms; falls back
    to polling nvidia-smi when pynvml is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag, self.n = index, [], [], set(), False, 0

    def _nvml(self):
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
        h = N.nvmlDeviceGetHandleByIndex(idx)
        self.mx.append(float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)))
        bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
            r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, b in bits.items():
                if r & b:
                    self.reasons.add(name)
            self.n += 1
            time.sleep(0.005)

    def _smi(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                r = [x.strip() for x in o.split(",")]
                if len(r) >= 6 and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0])); self.mx.append(float(r[1])); self.n += 1
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": self.n}


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def base_line(args, tasks):
    return {"metric": "Mbp polished/s", "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": "synthetic 5 Mb draft (5 x 1 Mb) + 30x 150 bp PE short reads per GPU; step = tasks %s" % tasks,
                       "tasks": tasks, "per_gpu_bp": WORKLOAD["n_contigs"] * WORKLOAD["contig_len"],
                       "parallelism": "contig-sharded x%d" % args.gpus, "task2_draft": TASK2_DRAFT}}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on all host cores
# ------------------------------------------------------------------------------------------------
class ReferenceRunner:
    """The compiled, unmodified reference binary (oracle/_ref/nextpolish1) on the files of rank 0's workload: one
    process per contig (the reference's parallel grain, nextpolish1.py:223-224) x `copies` concurrent copies of
    the step so that every host thread is busy.  No product code is loaded or executed."""

    def __init__(self, tmpdir, tasks):
        self.tasks, self.tmp = list(tasks), tmpdir
        self.ncpu = os.cpu_count() or 1
        self.files = write_inputs(tmpdir, 0, tasks)
        self.parts = {}
        for t in tasks:
            fa, bam = self.files[t]
            self.parts[t] = []
            for n, s in read_fasta_bytes(fa).items():
                f = os.path.join(tmpdir, "%s.t%d.fa" % (n, t))
                with open(f, "wb") as fh:
                    fh.write(b">" + n.encode() + b"\n" + s + b"\n")
                self.parts[t].append(f)
        self.n_sample = WORKLOAD["n_contigs"]          # contigs of every copy polished per step
        self.copies = max(1, self.ncpu // self.n_sample)

    def outputs_md5(self, task):
        """md5 of every contig the reference emits for `task` (one untimed pass, stdout captured)."""
        fa, bam = self.files[task]
        out = {}
        procs = [subprocess.Popen([REF_BIN, CMDS[task], f, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL) for f in self.parts[task]]
        for pr in procs:
            so, _ = pr.communicate()
            assert pr.returncode == 0
            name = None
            for line in so.split(b"\n"):
                if line.startswith(b">"):
                    name = line[1:].decode()
                    name = name[:name.rfind("_")]          # contig_write_to_file appends _<step> (contig.c:1050)
                elif name is not None and line:
                    out[name] = hashlib.md5(line.strip()).hexdigest()
        return out

    def one_step(self):
        for t in self.tasks:
            bam = self.files[t][1]
            procs = [subprocess.Popen([REF_BIN, CMDS[t], f, bam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                     for _ in range(self.copies) for f in self.parts[t][:self.n_sample]]
            for pr in procs:
                assert pr.wait() == 0

    def bound_sample(self, steps, warmup, budget_s):
        """Times one step, then shrinks the per-step sample (contigs per copy) so that steps+warmup fit the budget."""
        t0 = time.time()
        self.one_step()
        dt = time.time() - t0
        want = budget_s / max(1, steps + warmup)
        if dt > want:
            self.n_sample = max(1, min(WORKLOAD["n_contigs"], int(WORKLOAD["n_contigs"] * want / dt)))
            self.copies = max(1, self.ncpu // self.n_sample)

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.one_step()
        t0 = time.time()
        for _ in range(steps):
            self.one_step()
        dt = time.time() - t0
        bp_step = self.copies * self.n_sample * WORKLOAD["contig_len"] * len(self.tasks)
        sample = ("per step: %d concurrent copies x %d of the %d contigs (1 Mb, %gx) x tasks %s = %.0f Mbp; one process per contig "
                  "(the reference's parallel grain, nextpolish1.py:223-224), %d processes on %d host threads; BGZF BAM input; %d steps"
                  % (self.copies, self.n_sample, WORKLOAD["n_contigs"], WORKLOAD["depth"], self.tasks, bp_step / 1e6,
                     self.copies * self.n_sample, self.ncpu, steps))
        return bp_step * steps / dt / 1e6, dt / steps * 1e3, self.copies * self.n_sample, sample


_PORT_STATE = None


def _port_contig(args):
    import numpy as np
    c, task = args
    shs, cfg = _PORT_STATE
    sh = shs[task]
    O = C.CDLL(os.path.join(ROOT, "oracle", "libnp_oracle.so"))
    O.np_oracle_run_contig.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    cap = int(sh.view.ctg_off[c + 1] - sh.view.ctg_off[c]) * 2 + 4096
    out = np.zeros(cap, np.uint8)
    n = C.c_int64(0)
    assert O.np_oracle_run_contig(C.addressof(sh.view), c, task, C.cast(cfg, C.c_void_p), out.ctypes.data, cap, C.byref(n)) == 0
    return int(n.value)


def port_run(steps, warmup, tasks):
    """Fallback when the reference could not be compiled: the oracle port (oracle/np_oracle.c) on all host cores."""
    import multiprocessing as mp
    from nextpolish_b200 import engine as E
    global _PORT_STATE
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    _PORT_STATE = ({t: E.Shard.synthetic(E.synth_params(**synth_kwargs(t, seed_for(0, t))), 0, WORKLOAD["n_contigs"], with_qual=True) for t in tasks}, cfg)
    nproc = min(os.cpu_count() or 1, WORKLOAD["n_contigs"])
    pool = mp.get_context("fork").Pool(nproc)

    def one_step():
        for t in tasks:
            pool.map(_port_contig, [(c, t) for c in range(WORKLOAD["n_contigs"])])
    for _ in range(warmup):
        one_step()
    t0 = time.time()
    for _ in range(steps):
        one_step()
    dt = time.time() - t0
    pool.close()
    bp = WORKLOAD["n_contigs"] * WORKLOAD["contig_len"] * len(tasks)
    return bp * steps / dt / 1e6, dt / steps * 1e3, nproc, "oracle port, %d contigs x tasks %s per step, %d processes" % (WORKLOAD["n_contigs"], list(tasks), nproc)


def have_reference():
    return os.path.exists(REF_BIN) and os.path.exists(SAMTOOLS) and os.path.exists(SIMULATE)


def main_reference(args, tasks):
    base = base_line(args, tasks)
    with tempfile.TemporaryDirectory(prefix="npbench") as tmp:
        if have_reference():
            rr = ReferenceRunner(tmp, tasks)
            rr.bound_sample(args.steps, args.warmup, 240.0)
            v, ms, cores, sample = rr.run(max(1, args.steps), max(0, args.warmup))
            kind = "reference"
        else:
            v, ms, cores, sample = port_run(max(1, args.steps), max(0, args.warmup), tasks)
            kind = "port"
    base.update({"impl": "reference", "value": v, "ms_per_step": ms, "dtype": "int/f64 (CPU)",
                 "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample, "host_cpus": os.cpu_count()},
                 "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main_ours(args, tasks):
    rank, local_rank, world = rank_env()
    import numpy as np
    import torch
    import torch.distributed as dist
    from nextpolish_b200 import engine as E
    base = base_line(args, tasks)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tmp = tempfile.mkdtemp(prefix="npbench")
    # ---- inputs: N_ROTATE distinct shards of the workload shape, pinned on host and resident in HBM
    shards, pinned, resident, views_dev, views_host = {}, {}, {}, {}, {}
    for t in tasks:
        shards[t], pinned[t], resident[t], views_dev[t], views_host[t] = [], [], [], [], []
        for k in range(N_ROTATE):
            p = E.synth_params(**synth_kwargs(t, seed_for(rank, t, k)))
            sh = E.Shard.synthetic(p, 0, WORKLOAD["n_contigs"], with_qual=(2 if t == 2 else 0), threads=max(1, (os.cpu_count() or 8) // max(1, args.gpus)))
            a = sh.arrays()
            pin = {k2: torch.from_numpy(v.copy()).pin_memory() for k2, v in a.items() if k2 in ("ctg_seq", "rec_off", "rec", "qual_off", "qual")}
            res = {k2: x.to(dev) for k2, x in pin.items()}

            def mkview(src, sh=sh, t=t):
                v = E.ShardView()
                v.n_contigs, v.n_reads = sh.view.n_contigs, sh.view.n_reads
                v.ctg_off, v.ctg_read_off = sh.view.ctg_off, sh.view.ctg_read_off
                v.ctg_seq, v.rec_off, v.rec = src["ctg_seq"].data_ptr(), src["rec_off"].data_ptr(), src["rec"].data_ptr()
                if t == 2:
                    v.qual_off, v.qual = src["qual_off"].data_ptr(), src["qual"].data_ptr()
                return v
            shards[t].append(sh); pinned[t].append(pin); resident[t].append(res)
            views_dev[t].append(mkview(res)); views_host[t].append(mkview(pin))
    files = write_inputs(tmp, rank, tasks)          # the e2e inputs: the very files the reference arm reads (rank 0: same seed)
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750          # what config_init estimates on these BAMs (insert N(350,35) x 5)
    eng = E.Engine(local_rank)
    bp_step = sum(int(shards[t][0].total_bases) for t in tasks)      # bases polished per step on this rank
    alg_bytes = {t: shards[t][0].algorithmic_bytes(t) for t in tasks}
    h2d_packed = sum(sum(x.numel() * x.element_size() for x in pinned[t][0].values()) for t in tasks)
    cap = int(max(shards[t][0].total_bases for t in tasks) * 1.25) + 4096
    estream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    from nextpolish_b200.sharding import FixedGather
    GSLOTS = 2
    HDR = FixedGather.HEADER
    fixed_gather = FixedGather(cap, dev, slots=GSLOTS) if world > 1 else None
    gbufs = [torch.zeros(cap + HDR, dtype=torch.uint8, device=dev) for _ in range(GSLOTS)]
    gready = [torch.cuda.Event() for _ in range(GSLOTS)]
    gdone = [None] * GSLOTS
    gstate = {"job": 0}

    def gather_fasta():
        """The single collective of the path: corrected FASTA bytes of every rank -> rank 0 (one NCCL gather per
        task step; byte count in the buffer header).  Stream-ordered and double-buffered: the gather of one step
        runs on torch's NCCL stream while the engine stream already computes the next step; no host sync."""
        j = gstate["job"] % GSLOTS
        gstate["job"] += 1
        b = gbufs[j]
        if gdone[j] is not None:
            estream.wait_event(gdone[j])                     # the gather that last read this buffer is over
        # header (byte count) + polished bytes, written by the engine on its own stream from device memory: no torch
        # tensor op ever runs on the engine's stream, so torch's allocators hold no reference to it at shutdown
        rc = E.lib().np_engine_pack_result(eng.h, b.data_ptr(), cap + HDR)
        assert rc == 0, E.last_error()
        gready[j].record(estream)
        if world > 1:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(gready[j])
            fixed_gather(b, slot=j)
            ev = torch.cuda.Event()
            ev.record(cur)
            gdone[j] = ev

    def step_resident(i):
        for t in tasks:
            eng.adopt_device(views_dev[t][i % N_ROTATE])
            eng.run(t, cfg)
            gather_fasta()

    # resident shards through np_resident: RES_SLOTS engines (stream + scratch + host thread each) work on different
    # shards at once; slot = ticket % RES_SLOTS, one result buffer per slot; the gather of a finished job (N > 1) is
    # issued when its ticket is collected and must be over before the slot's buffer is written again
    rpipe = E.ResidentSlots(local_rank, RES_SLOTS)
    rbufs = [torch.zeros(cap + HDR, dtype=torch.uint8, device=dev) for _ in range(RES_SLOTS)]
    rgather = FixedGather(cap, dev, slots=RES_SLOTS) if world > 1 else None
    rdone = [None] * RES_SLOTS
    rpending = []
    rstate = {"next": 0}

    def resident_collect():
        tk = rpending.pop(0)
        rpipe.wait(tk)                                   # the job is complete on the device (host-synchronised)
        if world > 1:
            j = tk % RES_SLOTS
            rgather(rbufs[j], slot=j)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            rdone[j] = ev

    def step_resident_slots(i):
        for t in tasks:
            while len(rpending) >= RES_SLOTS:
                resident_collect()
            j = rstate["next"] % RES_SLOTS
            if rdone[j] is not None:
                rdone[j].synchronize()
                rdone[j] = None
            rpending.append(rpipe.submit(t, views_dev[t][i % N_ROTATE], cfg, rbufs[j].data_ptr(), cap + HDR))
            rstate["next"] += 1

    def flush_resident():
        while rpending:
            resident_collect()

    # e2e (packed): one output buffer per job in flight (a job's result lands in its own pinned buffer)
    DEPTH = 2
    pipe = E.Stream(local_rank, DEPTH)
    n_out = DEPTH + len(tasks)
    outs = [(torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(WORKLOAD["n_contigs"] + 1, dtype=torch.int64).pin_memory())
            for _ in range(n_out)]
    outs_np = [(a.numpy(), b.numpy()) for a, b in outs]
    pending = []
    state = {"job": 0, "d2h": 0}

    def step_packed(i):
        for t in tasks:
            o, f = outs_np[state["job"] % n_out]
            state["job"] += 1
            pending.append((pipe.submit(t, views_host[t][i % N_ROTATE], cfg, o, f), f))
            while len(pending) > DEPTH:          # results of the older jobs are read back (D2H) here
                tk, f0 = pending.pop(0)
                pipe.wait(tk)
                state["d2h"] = int(f0[-1]) + f0.nbytes

    def flush_packed():
        while pending:
            tk, f0 = pending.pop(0)
            pipe.wait(tk)
            state["d2h"] = int(f0[-1]) + f0.nbytes

    # e2e (files): FASTA + BAM (+ .bai) in the page cache -> polished bytes in host memory
    fpipe = E.FilePipeline(local_rank, depth=FILES_DEPTH)
    fstate = {"h2d": 0, "d2h": 0, "last": {}}

    # (the polished bytes of every job land in pinned host memory; hashing them for the parity check is not part of the
    # path: only the jobs of the final flush are hashed)
    def step_files(i):
        for t in tasks:
            fa, bam = files[t]
            fpipe.submit(t, fa, bam, cfg)
            while fpipe.in_flight() > fpipe.capacity - 1:
                r = fpipe.wait_oldest(want_md5=False)
                fstate["last"][r["task"]] = r
        return None

    def flush_files():
        while fpipe.in_flight():
            r = fpipe.wait_oldest(want_md5=False)
            fstate["last"][r["task"]] = r

    def timed(fn, steps, warmup, flush=None, split=None):
        for i in range(warmup):
            fn(i)
        if flush:
            flush()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(estream):
            e0.record()
        t0 = time.time()
        for i in range(steps):
            fn(i)
            if split is not None and (i + 1) % max(1, steps // 4) == 0:
                split.append(round((time.time() - t0) * 1e3, 1))    # wall ms at every quarter (diagnostics: drift inside the region)
        if flush:
            flush()                                            # every job finished and read back (host-synchronised)
        estream.wait_stream(torch.cuda.current_stream(dev))   # orders the (asynchronous) NCCL gathers before e1
        with torch.cuda.stream(estream):
            e1.record()
        barrier()
        wall = (time.time() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item()), float(ms[1].item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_serial, _ = timed(step_resident, args.steps, args.warmup)
    ms_res, _ = timed(step_resident_slots, args.steps, args.warmup + RES_SLOTS, flush_resident)
    # per-kernel times of one resident step (CUDA events on the engine stream)
    ktimes, kt_by_task = {}, {}
    launches_per_step = 0          # kernels launched by one resident step: counted per task run below
    wstats = None
    eng.set_timing(True)
    for t in tasks:
        eng.adopt_device(views_dev[t][0])
        eng.run(t, cfg)
        launches_per_step += eng.launch_count()
        if t == 1:
            wstats = eng.window_stats()
        eng.sync()
        kt = eng.kernel_times()
        kt_by_task[t] = kt
        for n, v in kt:
            ktimes[n] = ktimes.get(n, 0.0) + v
    eng.set_timing(False)
    ms_packed, _ = timed(step_packed, args.steps, args.warmup, flush_packed)
    # the files path keeps host threads busy (file reads, block scan): its wall clock is the honest number, the
    # device events bracket the same region
    e2e_steps = args.steps
    # the pipeline reaches its steady state only after every slot has run a few jobs (memory pools, pinned buffers, the
    # host's page tables for the mapped BAMs): warm up at least eight jobs per worker, untimed, then time exactly K steps
    files_warmup = max(args.warmup, 4 * FILES_DEPTH)
    files_split = []
    ms_files_dev, ms_files = timed(step_files, e2e_steps, files_warmup, flush_files, split=files_split)
    ms_files = max(ms_files, ms_files_dev)
    for t in tasks:                                  # one untimed job per task, hashed: what the parity check compares
        fpipe.submit(t, files[t][0], files[t][1], cfg)
        fstate["last"][t] = fpipe.wait_oldest(want_md5=True)
    files_out = {t: fstate["last"][t] for t in tasks}
    if sampler:
        sampler.stop_flag = True
        sampler.join()

    line = None
    if rank == 0:
        total_bp = bp_step * world
        value = total_bp * args.steps / (ms_res / 1e3) / 1e6
        e2e = total_bp * e2e_steps / (ms_files / 1e3) / 1e6
        e2e_packed = total_bp * args.steps / (ms_packed / 1e3) / 1e6
        peak, peak_kind = measured_peak_gbs()
        traffic = None
        try:   # DRAM bytes of one launch of the pileup-scan kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_pileup_diff_kernel.json")))
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        roof_kernel = "pileup_diff"
        k1 = dict(kt_by_task[1]) if 1 in kt_by_task else {}
        kms = k1.get(roof_kernel)
        pair_ms = (kms + k1["pileup_scan"]) if kms and "pileup_scan" in k1 else None
        t1_ms = sum(v for _, v in kt_by_task.get(1, []))
        ach = alg_bytes[1] / (kms / 1e3) / 1e9 if kms else None
        h2d_files = sum(files_out[t]["h2d_bytes"] for t in tasks)
        d2h_files = sum(files_out[t]["d2h_bytes"] for t in tasks)
        base.update({
            "value": value, "ms_per_step": ms_res / args.steps, "dtype": "u8/u16/int32 (+f64 score chain)",
            "resident": {"slots": RES_SLOTS, "api": "np_resident_submit/np_resident_wait: %d engines (stream + scratch + host thread each) "
                                                    "polish different resident shards concurrently" % RES_SLOTS,
                         "one_engine": {"value": total_bp * args.steps / (ms_serial / 1e3) / 1e6, "ms_per_step": ms_serial / args.steps,
                                        "what": "the same steps on ONE engine and stream, one task after the other (round 1's `value`); "
                                                "kernels_ms and the roofline are measured on this serial form"}},
            "e2e": {"value": e2e, "unit": "Mbp/s", "h2d_bytes_per_step": h2d_files, "d2h_bytes_per_step": d2h_files,
                    "ms_per_step": ms_files / e2e_steps, "ms_per_step_device_events": ms_files_dev / e2e_steps,
                    "warmup_steps": files_warmup, "wall_ms_at_quarters": files_split,
                    "api": "np_files_submit/np_files_wait (FASTA + BGZF BAM + .bai in the page cache -> polished bytes on the host; %d workers, up to %d jobs queued)" % (FILES_DEPTH, 2 * FILES_DEPTH)},
            "e2e_packed": {"value": e2e_packed, "unit": "Mbp/s", "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": state["d2h"] * len(tasks),
                           "ms_per_step": ms_packed / args.steps, "api": "np_stream_submit/np_stream_wait, pre-packed shards in pinned host memory, depth %d" % DEPTH},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None, "traffic": traffic, "peak_kind": peak_kind,
                         "algorithmic_bytes": alg_bytes[1], "kernel_ms": kms,
                         "pair": {"kernels_ms": pair_ms, "achieved": alg_bytes[1] / (pair_ms / 1e3) / 1e9 if pair_ms else None,
                                  "frac": alg_bytes[1] / (pair_ms / 1e3) / 1e9 / peak if pair_ms else None,
                                  "what": "the same algorithmic bytes over pileup_diff + pileup_scan (diff pass + column pass)"},
                         "task1": {"kernels_ms": t1_ms, "achieved": alg_bytes[1] / (t1_ms / 1e3) / 1e9 if t1_ms else None,
                                   "frac": alg_bytes[1] / (t1_ms / 1e3) / 1e9 / peak if t1_ms else None,
                                   "what": "the same algorithmic bytes over the summed device time of every task-1 kernel"}},
            "kernels_ms": {k: round(v, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])},
            "kernels_ms_by_task": {str(t): round(sum(v for _, v in kt_by_task[t]), 4) for t in kt_by_task},
            "pileup_windows": wstats,
            "clocks": sampler.summary() if sampler else None,
        })
        base["config"].update({"l2": "inputs rotate over %d distinct resident shards (%.0f MB each) and every step rewrites "
                                     ">300 MB of scratch: working set exceeds the 126 MB L2" % (N_ROTATE, h2d_packed / 1e6),
                               "reads_per_gpu": int(shards[1][0].n_reads), "algorithmic_bytes_per_bp": alg_bytes[1] / shards[1][0].total_bases})
        if args.gpus == 1 and not args.no_cpu_baseline:
            try:
                if have_reference():
                    rr = ReferenceRunner(tmp, tasks)       # rewrites rank 0's files (same seeds -> same bytes)
                    ref_md5 = {t: rr.outputs_md5(t) for t in tasks}
                    ours_md5 = {t: files_out[t]["md5"] for t in tasks}
                    base["parity_in_bench"] = all(ref_md5[t] == ours_md5[t] and len(ref_md5[t]) == WORKLOAD["n_contigs"] for t in tasks)
                    base["parity_in_bench_what"] = ("md5 of every polished contig of the e2e (files) path vs the reference binary's stdout "
                                                    "on the same FASTA + BAM, tasks %s: %d contigs compared" % (tasks, sum(len(ref_md5[t]) for t in tasks)))
                    rr.bound_sample(2, 0, 30.0)
                    v, ms, cores, sample = rr.run(2, 0)
                    kind = "reference"
                else:
                    v, ms, cores, sample = port_run(2, 1, tasks)
                    kind = "port"
                base["cpu_baseline"] = {"value": v, "unit": "Mbp/s", "cores": cores, "kind": kind, "sample": sample, "host_cpus": os.cpu_count()}
            except Exception as ex:          # informational keys: never lose the bench line over them
                base["cpu_baseline"] = {"error": repr(ex)}
            try:
                base["task2_sensitivity"] = task2_dense_run(E, eng, cfg, dev, args)
            except Exception as ex:
                base["task2_sensitivity"] = {"error": repr(ex)}
        line = json.dumps(base)

    if line:
        print(line)
        sys.stdout.flush()
    # ordered shutdown: engines / pipelines (their CUDA streams) first, then the process group; the interpreter then
    # exits normally (no os._exit) so that exit hooks run
    torch.cuda.synchronize()
    rpipe.close()
    fpipe.close()
    pipe.close()
    eng.close()
    for t in tasks:
        for sh in shards[t]:
            sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    shutil.rmtree(tmp, ignore_errors=True)


def task2_dense_run(E, eng, cfg, dev, args):
    """Task 2 alone on a draft with the lowercase density of real data (4.0e-3): resident Mbp/s next to the bench's
    own task-2 input (6.3e-4)."""
    import torch
    out = {}
    for name, dense in (("bench_6.3e-4", False), ("dense_4.0e-3", True)):
        p = E.synth_params(**synth_kwargs(2, seed_for(0, 2, 7), dense=dense))
        sh = E.Shard.synthetic(p, 0, WORKLOAD["n_contigs"], with_qual=2, threads=os.cpu_count() or 8)
        eng.upload(sh.view)
        for _ in range(3):
            eng.run(2, cfg)
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.ExternalStream(eng.stream(), device=dev)
        n = 20
        with torch.cuda.stream(s):
            e0.record()
        for _ in range(n):
            eng.run(2, cfg)
        with torch.cuda.stream(s):
            e1.record()
        eng.sync()
        ms = e0.elapsed_time(e1) / n
        out[name] = {"task2_ms": ms, "Mbp_per_s": int(sh.total_bases) / ms / 1e3}
        sh.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, _, _ = rank_env()
    tasks = [1, 2]
    if args.impl == "reference":
        if rank == 0:
            main_reference(args, tasks)
        return
    main_ours(args, tasks)


if __name__ == "__main__":
    main()
