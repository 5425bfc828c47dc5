# round 2, session a: tests with the overlap protocol, bench by overlap mode, the full default line
exec > gpurun_out/session_r2a.log 2>&1
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for ov in 0 1 2; do
  timeout 300 python bench.py --no-extras --overlap $ov --steps 20 --warmup 5 --cpu-seconds 1 --e2e-steps 20 > gpurun_out/bench_r2a_ov$ov.json 2> gpurun_out/bench_r2a_ov$ov.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_r2a_ov$ov.json'));print('OV$ov', d['ms_per_step']*1e3,'us frac',d['roofline']['frac'],'e2e us',d['e2e']['ms_per_step']*1e3, d['repeats'])"
done
timeout 300 python bench.py --no-extras --overlap 2 --steps 240 --warmup 8 --cpu-seconds 1 --e2e-steps 20 > gpurun_out/bench_r2a_ov2_k240.json 2>> gpurun_out/bench_r2a_ov2.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a_full.json 2> gpurun_out/bench_r2a_full.err
tail -5 gpurun_out/bench_r2a_full.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err
cut -c1-600 gpurun_out/bench_r2a_ref.json
