mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "oracle_on_synthetic or general_kernels or deep_and_noisy" > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
for W in 512 256 128; do for T in 64 128 256; do
  NEXTPOLISH_B200_WINDOW=$W NEXTPOLISH_B200_WIN_THREADS=$T python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('W=$W T=$T', 'pileup_ms', d['kernels_ms'].get('pileup_scan'), 'smem', d['pileup_windows']['smem_bytes'], 'value', round(d['value']))"
done; done
