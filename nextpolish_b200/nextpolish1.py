#!/usr/bin/env python
"""Host-side mirror of the reference's per-step worker (source/lib/nextpolish1.py): same command line,
same block-file / resume / FASTA-header conventions, but every unpolished contig of the block is
polished in ONE batch on the GPU through the np_* C ABI instead of one ctypes call per contig in a
multiprocessing.Pool.

    python -m nextpolish_b200.nextpolish1 -g genome.fa -t 1 -s sgs.sort.bam -o genome.nextpolish.part000.fasta
                                          [-b input.genome.fasta.blc -i 0] [-u] [algorithm flags]

Behavioural contract kept from the reference (nextpolish1.py:148-235):
  * -b/-i select the contigs of block i ("name<TAB>index" lines); default: every contig of -g;
  * an existing -o is scanned, finished contigs are skipped and the file is truncated at its last
    (possibly partial) record;
  * records are written as ">name_np<task> <len>\\n<seq>" (a name already ending in _np... gets the task
    digit appended); -u uppercases;
  * Configure fields are set after config_init exactly like update_cfg() does (so read_tlen keeps the
    estimate made with the default count_read_ins_sgs / max_ins_len_sgs / max_ins_fold_sgs).
  * -debug sets Configure.trace_polish_open and prints "name pos index curbase base" change points to stderr
    (nextpolish1.py:133,230-231).
The shard is built on the GPU from the BAM's compressed bytes when <bam>.bai exists (np_shard_load_gpu), by the host
packer otherwise.  Tasks 1 (score_chain), 2 (kmer_count) and 4 (snp_valid) run on the GPU; not supported: tasks 3 and 5
(exit code 1, like the reference does for task 5)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.realpath(__file__))))


def read_block(path, index, polished):
    """Contigs to polish: block-file lines "name<TAB>index" with that index (nextpolish1.py:148-161), or
    every FASTA header when index == "all"; names already present in the output are dropped."""
    names = []
    with open(path) as f:
        for line in f:
            if index == "all":
                if line.startswith(">"):
                    names.append(line.strip().split()[0][1:])
            else:
                parts = line.strip().split()
                if len(parts) >= 2 and parts[1] == index:
                    names.append(parts[0])
    return [n for n in names if n.split("_np")[0] not in polished]


def scan_output(path, polished):
    """Returns the byte offset at which the last (possibly partial) record starts."""
    last, offset, cur = "", 0, 0
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                offset += cur
                cur = len(line)
                last = line.split()[0].split("_np")[0][1:]
                polished.add(last)
            else:
                cur += len(line)
    if last:
        polished.discard(last)
    return offset


def main(argv=None):
    ap = argparse.ArgumentParser(description="Polish the genome on the GPU (drop-in for lib/nextpolish1.py).")
    ap.add_argument("-g", "--genome", required=True)
    ap.add_argument("-s", "--bam_sgs")
    ap.add_argument("-l", "--bam_lgs")
    ap.add_argument("-b", "--block")
    ap.add_argument("-i", "--block_index", default="all")
    ap.add_argument("-u", "--uppercase", action="store_true")
    ap.add_argument("-debug", action="store_true")
    ap.add_argument("-o", "--out", default="stdout")
    ap.add_argument("-t", "--task", type=int, required=True, choices=[1, 2, 3, 4, 5])
    ap.add_argument("-p", "--process", type=int, default=10, help="host threads for BAM decoding")
    for flag, typ, dflt in (("count_read_ins_sgs", int, 10000), ("min_map_quality", int, 0), ("max_ins_len_sgs", int, 10000),
                            ("max_ins_fold_sgs", int, 5), ("max_clip_ratio_sgs", float, 0.15), ("max_clip_ratio_lgs", float, 0.4),
                            ("trim_len_edge", int, 2), ("ext_len_edge", int, 2), ("indel_balance_factor_sgs", float, 0.5),
                            ("min_count_ratio_skip", float, 0.8), ("min_len_ldr", int, 3), ("max_len_kmer", int, 50),
                            ("min_len_inter_kmer", int, 5), ("max_count_kmer", int, 50)):
        ap.add_argument("-" + flag, type=typ, default=dflt)
    args, _unknown = ap.parse_known_args(argv)
    if args.task not in (1, 2, 4):
        sys.stderr.write("tasks 3 and 5 are outside this engine's scope: use the reference nextpolish1.py / nextpolish2.py\n")
        return 1
    from nextpolish_b200 import engine as E

    out, polished = sys.stdout, set()
    if args.out != "stdout":
        if os.path.exists(args.out):
            pos = scan_output(args.out, polished)
            out = open(args.out, "r+")
            out.seek(pos)
            out.truncate()
        else:
            out = open(args.out, "w")
    block = args.genome if (args.block_index == "all" or not args.block) else args.block
    index = "all" if block == args.genome else args.block_index
    names = read_block(block, index, polished)

    cfg = E.default_config(args.genome, args.bam_sgs)
    c = cfg.contents                                    # update_cfg (nextpolish1.py:102-133)
    for f in ("trim_len_edge", "ext_len_edge", "min_map_quality", "indel_balance_factor_sgs", "min_count_ratio_skip", "min_len_ldr",
              "min_len_inter_kmer", "max_len_kmer", "max_count_kmer", "count_read_ins_sgs", "max_ins_len_sgs", "max_ins_fold_sgs",
              "max_clip_ratio_sgs", "max_clip_ratio_lgs"):
        setattr(c, f, getattr(args, f))
    c.trace_polish_open = 1 if args.debug else 0        # nextpolish1.py:133
    if names:
        dev = int(os.environ.get("NEXTPOLISH_B200_DEVICE", "0"))
        wq = 2 if args.task == 2 else 1 if args.task == 4 else 0
        eng = E.Engine(dev)
        shard = None
        if args.bam_sgs and os.path.exists(args.bam_sgs + ".bai") and os.environ.get("NEXTPOLISH_B200_HOST_LOAD") != "1":
            try:
                shard = E.DeviceShard(args.genome, args.bam_sgs, names=names, with_qual=wq, device=dev)
            except E.NativeError:
                shard = None                             # e.g. a stale index: the host packer takes over
        if shard is not None:
            eng.adopt_device(shard.view)
            eng.run(args.task, cfg)
            raw, off = eng.download(shard.n_contigs)
            raw = raw.tobytes()
            seqs = {nm: raw[off[i]:off[i + 1]] for i, nm in enumerate(shard.names)}
        else:
            shard = E.Shard.load(args.genome, args.bam_sgs, names=names, with_qual=wq, threads=max(1, args.process))
            seqs = eng.polish(shard, args.task, cfg)
        points = dict(zip(shard.names, eng.points(shard.n_contigs))) if args.debug else {}
        for name in names:                               # the reference's order is completion order; ours is block order
            seq = seqs[name].decode()
            if args.uppercase:
                seq = seq.upper()
            tag = name + (str(args.task) if name.split("_")[-1].startswith("np") else "_np" + str(args.task))
            out.write(">%s %d\n%s\n" % (tag, len(seq), seq))
            for pos, idx, cur, base in points.get(name, ()):
                sys.stderr.write("%s %d %d %s %s\n" % (name, pos, idx, cur.decode(), base.decode()))
        shard.close()
        eng.close()
    if out is not sys.stdout:
        out.close()
    E.lib().config_destory(cfg)
    return 0


if __name__ == "__main__":
    sys.exit(main())
