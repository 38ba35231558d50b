"""np_multi (csrc/multi_gpu.cu): one input on several GPUs of one box — contiguous contig blocks, one NCCL gather.
The single-GPU form runs everywhere; the 2-GPU form needs a box with two GPUs (gpurun --gpus 2)."""
import os
import subprocess

import pytest

from tests.conftest import GOLDEN, REF_SAMTOOLS, read_fasta, run_checker

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_gpus", [1, 2])
@pytest.mark.parametrize("task", [1, 2, 4])
def test_multi_gpu_equals_oracle(E, oracle, synth_files, n_gpus, task):
    if _ngpu() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    fa, bam = synth_files("ragged")              # 24 contigs of ragged length: blocks of several contigs per GPU
    if not os.path.exists(bam + ".bai"):
        subprocess.check_call([REF_SAMTOOLS, "index", bam])
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    want = run_checker(oracle.np_oracle_run, sh, task, cfg)
    m = E.MultiGpu(n_gpus)
    got, stats = m.polish(task, fa, bam, cfg)
    got2, _ = m.polish(task, fa, bam, cfg)       # buffers are reused by the next run
    m.close()
    assert list(got) == list(read_fasta(fa))     # FASTA order
    assert got == want and got2 == want
    assert stats["d2h_bytes"] == sum(len(s) for s in want.values())


def test_multi_gpu_many_blocks_under_a_small_block_budget(E, oracle, synth_files, monkeypatch):
    """A draft larger than the block budget is polished block by block (NEXTPOLISH_B200_BLOCK_MBP): same bytes."""
    fa, bam = synth_files("ragged")
    if not os.path.exists(bam + ".bai"):
        subprocess.check_call([REF_SAMTOOLS, "index", bam])
    sh = E.Shard.load(fa, bam, with_qual=True)
    cfg = E.default_config(fa, bam)
    want = run_checker(oracle.np_oracle_run, sh, 1, cfg)
    monkeypatch.setenv("NEXTPOLISH_B200_BLOCK_MBP", "0.02")        # 20 kb per block: ~10 blocks for the 216 kb ragged set
    n = min(2, _ngpu())
    m = E.MultiGpu(n)
    got, stats = m.polish(1, fa, bam, cfg)
    m.close()
    assert got == want


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_native_cli_on_several_gpus_matches_reference_output(E, n_gpus):
    if _ngpu() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    cli = os.path.join(os.path.dirname(E.binding.LIB_PATH), "nextpolish1")
    env = dict(os.environ, NEXTPOLISH_B200_GPUS=str(n_gpus))
    for step, cmd in ((1, "scorechain"), (2, "kmercount")):
        fa = os.path.join(GOLDEN, "td30.step%d.fa" % step)
        bam = os.path.join(GOLDEN, "td30.step%d.bam" % step)
        ours = subprocess.run([cli, cmd, fa, bam], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True, env=env).stdout
        assert ours == open(os.path.join(GOLDEN, "td30.step%d.expected.fa" % step), "rb").read(), cmd
