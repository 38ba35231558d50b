"""Python-side handle classes over the C ABI (host plumbing only; all compute is in the .so).

    lib   = binding.load()
    shard = Shard.load(fasta, bam, with_qual=False)          # FASTA + BAM -> packed shard
    shard = Shard.synthetic(params, 0, n_contigs)            # seeded synthetic shard
    eng   = Engine(device=0)
    seqs  = eng.polish(shard, task=1, cfg=cfg)               # {name: polished sequence bytes}
"""
import ctypes as C

import numpy as np

from . import binding  # noqa: F401  (re-exported: engine.binding)
from .binding import Configure, ShardView, SynthParams

_LIB = None
TASKS = (1, 2)  # task steps the device engine implements (1 = score_chain, 2 = kmer_count)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = binding.load()
    return _LIB


def last_error():
    return lib().np_last_error().decode(errors="replace")


class NativeError(RuntimeError):
    pass


def default_config(fasta=b"", bam=None):
    """Configure with the reference defaults (config.c:11-40); read_tlen from the BAM head."""
    return lib().config_init(fasta if isinstance(fasta, bytes) else fasta.encode(),
                             None if bam is None else (bam if isinstance(bam, bytes) else bam.encode()), None)


def synth_params(seed=1, n_contigs=1, contig_len=100000, depth=30.0, read_len=150, min_len=0, max_len=0,
                 draft_snv=0.001, draft_indel=0.003, read_sub=0.002, read_indel=0.0001,
                 lowercase_frac=0.0, compress_level=1):
    p = SynthParams()
    p.seed, p.n_contigs, p.contig_len, p.min_len, p.max_len = seed, n_contigs, contig_len, min_len, max_len
    p.depth, p.read_len = depth, read_len
    p.draft_snv, p.draft_indel, p.read_sub, p.read_indel = draft_snv, draft_indel, read_sub, read_indel
    p.lowercase_frac, p.compress_level = lowercase_frac, compress_level
    return p


class Shard:
    """Host-owned packed shard."""

    def __init__(self, handle):
        if not handle:
            raise NativeError(last_error())
        self.h = handle
        self.view = ShardView()
        lib().np_shard_view_of(self.h, C.byref(self.view))
        self.names = [lib().np_shard_contig_name(self.h, i).decode() for i in range(self.view.n_contigs)]

    @classmethod
    def load(cls, fasta, bam, names=None, with_qual=False, threads=8):
        """with_qual: False/0 none, True/1 every read, 2 only reads overlapping lowercase draft bases."""
        arr, n = None, 0
        if names:
            n = len(names)
            arr = (C.c_char_p * n)(*[s.encode() for s in names])
        return cls(lib().np_shard_load(fasta.encode(), bam.encode() if bam else None, arr, n, int(with_qual), threads))

    @classmethod
    def synthetic(cls, params, lo, hi, with_qual=False, threads=8):
        return cls(lib().np_synth_shard(C.byref(params), lo, hi, int(with_qual), threads))

    @property
    def n_contigs(self):
        return self.view.n_contigs

    @property
    def n_reads(self):
        return self.view.n_reads

    @property
    def total_bases(self):
        return self.view.ctg_off[self.view.n_contigs]

    def algorithmic_bytes(self, task):
        return lib().np_shard_algorithmic_bytes(self.h, task)

    def arrays(self):
        """numpy views (no copy) of the packed arrays: dict name -> ndarray."""
        v = self.view
        n, r = v.n_contigs, v.n_reads

        def arr(ptr, count, dt):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dt)
            addr = ptr if isinstance(ptr, int) else C.addressof(ptr.contents)
            buf = (C.c_uint8 * (count * np.dtype(dt).itemsize)).from_address(addr)
            return np.frombuffer(buf, dtype=dt, count=count)

        out = {
            "ctg_off": arr(v.ctg_off, n + 1, np.int64),
            "ctg_read_off": arr(v.ctg_read_off, n + 1, np.int64),
            "rec_off": arr(v.rec_off, r + 1, np.uint32),
        }
        out["ctg_seq"] = arr(v.ctg_seq, int(out["ctg_off"][-1]), np.uint8)
        out["rec"] = arr(v.rec, int(out["rec_off"][-1]) * 16, np.uint8)
        if v.qual_off:
            out["qual_off"] = arr(v.qual_off, r + 1, np.uint32)
            out["qual"] = arr(v.qual, int(out["qual_off"][-1]) * 16, np.uint8)
        return out

    def close(self):
        if self.h:
            lib().np_shard_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceShard:
    """Packed shard built in HBM from FASTA + BAM (+ .bai) by np_shard_load_gpu (inflate, record unpack and packing
    on the GPU).  `view` holds device pointers: pass it to Engine.adopt_device."""

    def __init__(self, fasta, bam, names=None, with_qual=False, device=0):
        arr, n = None, 0
        if names:
            n = len(names)
            arr = (C.c_char_p * n)(*[s.encode() for s in names])
        self.h = lib().np_shard_load_gpu(device, fasta.encode(), bam.encode(), arr, n, int(with_qual))
        if not self.h:
            raise NativeError(last_error())
        self.view = ShardView()
        lib().np_dev_shard_view(self.h, C.byref(self.view))
        self.names = [lib().np_dev_shard_contig_name(self.h, i).decode() for i in range(self.view.n_contigs)]

    @property
    def n_contigs(self):
        return self.view.n_contigs

    @property
    def n_reads(self):
        return self.view.n_reads

    @property
    def total_bases(self):
        return self.view.ctg_off[self.view.n_contigs]

    def stats(self):
        sizes = (C.c_int64 * 5)()
        ms = (C.c_float * 2)()
        lib().np_dev_shard_stats(self.h, sizes, ms)
        return dict(rec_bytes=sizes[0], qual_bytes=sizes[1], draft_bytes=sizes[2], compressed_bytes=sizes[3],
                    inflated_bytes=sizes[4], inflate_kernel_ms=ms[0], device_ms=ms[1])

    def arrays(self):
        """Host copies of the device arrays (tests): same keys as Shard.arrays()."""
        st = self.stats()
        n, r = self.view.n_contigs, self.view.n_reads
        out = {"ctg_off": np.array([self.view.ctg_off[i] for i in range(n + 1)], np.int64),
               "ctg_read_off": np.array([self.view.ctg_read_off[i] for i in range(n + 1)], np.int64),
               "rec_off": np.zeros(r + 1, np.uint32), "ctg_seq": np.zeros(st["draft_bytes"], np.uint8),
               "rec": np.zeros(st["rec_bytes"], np.uint8)}
        has_q = bool(self.view.qual_off)
        if has_q:
            out["qual_off"] = np.zeros(r + 1, np.uint32)
            out["qual"] = np.zeros(st["qual_bytes"], np.uint8)
        rc = lib().np_dev_shard_download(self.h, out["ctg_seq"].ctypes.data, out["rec_off"].ctypes.data, out["rec"].ctypes.data,
                                         out["qual_off"].ctypes.data if has_q else None, out["qual"].ctypes.data if has_q else None)
        if rc != 0:
            raise NativeError(last_error())
        return out

    def close(self):
        if self.h:
            lib().np_dev_shard_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ResidentSlots:
    """np_resident: `slots` engines (stream + scratch + host thread each) polishing shards that already sit in HBM,
    concurrently.  submit() hands the job to slot ticket % slots (wait for ticket - slots first); the result lands in
    the caller's device buffer as a 16-byte header (byte count) followed by the polished bytes."""

    def __init__(self, device=0, slots=4):
        self.slots = slots
        self.h = lib().np_resident_create(device, slots)
        if not self.h:
            raise NativeError(last_error())

    def submit(self, task, dev_view, cfg, dst_ptr, dst_cap):
        t = lib().np_resident_submit(self.h, task, C.byref(dev_view), cfg, dst_ptr, dst_cap)
        if t < 0:
            raise NativeError("rc=%d: %s" % (t, last_error()))
        return t

    def wait(self, ticket):
        n = C.c_int64(0)
        rc = lib().np_resident_wait(self.h, ticket, C.byref(n))
        if rc != 0:
            raise NativeError("rc=%d: %s" % (rc, last_error()))
        return n.value

    def launch_count(self):
        return lib().np_resident_launch_count(self.h)

    def close(self):
        if self.h:
            lib().np_resident_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stream:
    """np_stream: jobs (task, host shard) submitted in order; a job's upload overlaps the kernels of the jobs
    before it (double buffering).  Buffers handed to submit() must stay alive until wait(ticket) returned."""

    def __init__(self, device=0, depth=2):
        self.h = lib().np_stream_create(device, depth)
        if not self.h:
            raise NativeError(last_error())

    def submit(self, task, view, cfg, out, off):
        t = lib().np_stream_submit(self.h, task, C.byref(view), cfg, out.ctypes.data, out.size, off.ctypes.data)
        if t < 0:
            raise NativeError("rc=%d: %s" % (t, last_error()))
        return t

    def wait(self, ticket):
        rc = lib().np_stream_wait(self.h, ticket)
        if rc != 0:
            raise NativeError("rc=%d: %s" % (rc, last_error()))

    def launch_count(self):
        return lib().np_stream_launch_count(self.h)

    def close(self):
        if self.h:
            lib().np_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpu:
    """np_multi: ONE input (draft FASTA + BAM) polished on several GPUs of this box — contiguous contig blocks, a few
    pipelined slots per GPU, one NCCL gather of the polished bytes to the first GPU at the end (csrc/multi_gpu.cu)."""

    def __init__(self, n_gpus, devices=None):
        arr = (C.c_int32 * n_gpus)(*devices) if devices else None
        self.h = lib().np_multi_create(arr, n_gpus)
        if not self.h:
            raise NativeError(last_error())

    def polish(self, task, fasta, bam, cfg, names=None):
        """-> ({name: polished bytes} in FASTA order, stats dict); names: only these contigs (a worker's block)"""
        from .binding import FilesResult
        r = FilesResult()
        if names is None:
            rc = lib().np_multi_run(self.h, task, fasta.encode(), bam.encode(), cfg, C.byref(r))
        else:
            arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
            rc = lib().np_multi_run_names(self.h, task, fasta.encode(), bam.encode(), cfg, arr, len(names), C.byref(r))
        if rc != 0:
            raise NativeError("rc=%d: %s" % (rc, last_error()))
        seqs = {r.names[i].decode(): C.string_at(r.seq + r.start[i], r.len[i]) for i in range(r.n_contigs)}
        return seqs, {"h2d_bytes": r.h2d_bytes, "d2h_bytes": r.d2h_bytes, "blocks_ms": r.load_ms, "gather_download_ms": r.polish_ms}

    def close(self):
        if self.h:
            lib().np_multi_destroy(self.h)
            self.h = None


def partition_contiguous(lengths, n_parts):
    """np_partition_contiguous: block index of every contig (contiguous blocks, balanced cumulative length)."""
    import numpy as np
    a = np.asarray(lengths, np.int64)
    out = np.zeros(len(a), np.int32)
    lib().np_partition_contiguous(a.ctypes.data, len(a), n_parts, out.ctypes.data)
    return out.tolist()


class FilePipeline:
    """np_files: FASTA + BAM (+ .bai) files -> polished sequences on the host; `depth` workers (engine + host thread each)
    pull jobs from a queue of up to `capacity` = 2 x depth submitted, uncollected jobs (load, upload, inflate and unpack of
    one job overlap the kernels of the others)."""

    def __init__(self, device=0, depth=2):
        self.capacity = 2 * max(1, min(8, depth))
        self.h = lib().np_files_create(device, depth)
        if not self.h:
            raise NativeError(last_error())
        self.tickets = []

    def submit(self, task, fasta, bam, cfg):
        t = lib().np_files_submit(self.h, task, fasta.encode(), bam.encode(), cfg)
        if t < 0:
            raise NativeError("rc=%d: %s" % (t, last_error()))
        self.tickets.append(t)
        return t

    def in_flight(self):
        return len(self.tickets)

    def wait_oldest(self, want_seqs=False, want_md5=True):
        """Finishes the oldest outstanding job -> dict(task, names, md5 {name: hex}, h2d_bytes, d2h_bytes[, seqs]).
        want_md5=False skips the hashing (the result bytes stay valid in the slot until its next submit)."""
        import hashlib
        from .binding import FilesResult
        t = self.tickets.pop(0)
        r = FilesResult()
        rc = lib().np_files_wait(self.h, t, C.byref(r))
        if rc != 0:
            raise NativeError("rc=%d: %s" % (rc, last_error()))
        names = [r.names[i].decode() for i in range(r.n_contigs)]
        out = {"task": r.task, "names": names, "h2d_bytes": r.h2d_bytes, "d2h_bytes": r.d2h_bytes,
               "load_ms": r.load_ms, "polish_ms": r.polish_ms, "md5": {}}
        seqs = {}
        for i, nm in enumerate(names):
            if not (want_md5 or want_seqs):
                break
            b = C.string_at(r.seq + r.start[i], r.len[i])
            if want_md5:
                out["md5"][nm] = hashlib.md5(b).hexdigest()
            if want_seqs:
                seqs[nm] = b
        if want_seqs:
            out["seqs"] = seqs
        return out

    def close(self):
        if self.h:
            lib().np_files_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One GPU's polishing engine (np_engine). Creating it fails loudly without a CUDA device."""

    def __init__(self, device=0):
        self.h = lib().np_engine_create(device)
        if not self.h:
            raise NativeError(last_error())

    def _check(self, rc):
        if rc != 0:
            raise NativeError("rc=%d: %s" % (rc, last_error()))

    def upload(self, view):
        self._check(lib().np_engine_upload(self.h, C.byref(view)))

    def adopt_device(self, view):
        self._check(lib().np_engine_adopt_device(self.h, C.byref(view)))

    def run(self, task, cfg):
        self._check(lib().np_engine_run(self.h, task, cfg))

    def sync(self):
        self._check(lib().np_engine_sync(self.h))

    def result_bytes(self):
        return lib().np_engine_result_bytes(self.h)

    def download(self, n_contigs):
        n = self.result_bytes()
        out = np.empty(n + 1, dtype=np.uint8)
        off = np.empty(n_contigs + 1, dtype=np.int64)
        self._check(lib().np_engine_download(self.h, out.ctypes.data, n + 1, off.ctypes.data))
        return out[:n], off

    def polish_host(self, task, view, cfg, out, off):
        """np_polish_host: upload + run + download into caller-provided numpy buffers."""
        self._check(lib().np_polish_host(self.h, task, C.byref(view), cfg, out.ctypes.data, out.size, off.ctypes.data))

    def polish(self, shard, task, cfg):
        self.upload(shard.view)
        self.run(task, cfg)
        out, off = self.download(shard.n_contigs)
        raw = out.tobytes()
        return {nm: raw[off[i]:off[i + 1]] for i, nm in enumerate(shard.names)}

    def points(self, n_contigs):
        """PolishPoint trace of the last run (cfg.contents.trace_polish_open = 1): list per contig of
        (pos, index, curbase, base) tuples."""
        from .binding import PolishPoint
        n = lib().np_engine_point_count(self.h)
        buf = (PolishPoint * max(int(n), 1))()
        off = np.zeros(n_contigs + 1, dtype=np.int64)
        self._check(lib().np_engine_points(self.h, buf, max(int(n), 0), off.ctypes.data))
        return [[(buf[i].pos, buf[i].index, buf[i].curbase, buf[i].base) for i in range(off[k], off[k + 1])] for k in range(n_contigs)]

    def set_timing(self, on):
        lib().np_engine_set_timing(self.h, int(on))

    def kernel_times(self):
        cap = 256
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = lib().np_engine_kernel_times(self.h, names, ms, cap)
        return [(names[i].decode(), ms[i]) for i in range(n)]

    def window_stats(self):
        a = (C.c_int32 * 5)()
        lib().np_engine_window_stats(self.h, a)
        return dict(zip(("W", "n_win", "smem_bytes", "unresolved_windows", "fallback_cols"), list(a)))

    def launch_count(self):
        return lib().np_engine_launch_count(self.h)

    def stream(self):
        return lib().np_engine_stream(self.h)

    def close(self):
        if self.h:
            lib().np_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
