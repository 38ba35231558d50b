mkdir -p gpurun_out
for W in ${SWEEP_W:-512 1024}; do for T in ${SWEEP_T:-128 192 256}; do
  NEXTPOLISH_B200_WINDOW=$W NEXTPOLISH_B200_WIN_THREADS=$T python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('W=$W T=$T', 'pileup_ms', d['kernels_ms'].get('pileup_scan'), 'smem', d['pileup_windows']['smem_bytes'], 'fallback', d['pileup_windows']['fallback_cols'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
