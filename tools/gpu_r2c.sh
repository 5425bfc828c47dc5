# round 2, session c: tests after the trigger fix, dense variant timing, ncu captures (65 536 both variants, 1 M)
exec > gpurun_out/session_r2c.log 2>&1
set -x
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^Environment init" | tail -40
for ov in 0 2; do
  timeout 300 python bench.py --no-extras --overlap $ov --steps 20 --warmup 5 --cpu-seconds 1 --e2e-steps 20 > gpurun_out/bench_r2c_ov$ov.json 2> gpurun_out/bench_r2c_ov$ov.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_r2c_ov$ov.json'));print('OV$ov', d['ms_per_step']*1e3,'us frac',d['roofline']['frac'],'e2e us',d['e2e']['ms_per_step']*1e3, d['repeats'])"
done
T="timeout 300 python tools/step_timing.py"
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 65536 --worlds 1 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $T --task vss --envs 65536 --worlds 1 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 65536 --worlds 2 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 32768 --worlds 16 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 16384 --worlds 32 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 8192 --worlds 64 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $T --task vss --envs 4096 --worlds 128 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $T --task vss --envs 8192 --worlds 64 --steps 6000
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $T --task vss --envs 4096 --worlds 128 --steps 6000
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_vss_env_step --launch-skip 1210 -o gpurun_out/prof_r2c_vss65536_serial python tools/step_timing.py --task vss --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2c_a.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $N -k regex:k_vss_env_step --launch-skip 1210 -o gpurun_out/prof_r2c_vss65536_dense python tools/step_timing.py --task vss --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2c_b.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=2 $N -k regex:k_vss_env_step --launch-skip 604 -o gpurun_out/prof_r2c_vss1m_dense python tools/step_timing.py --task vss --envs 1048576 --worlds 2 --warmup 300 --no-graph --steps 8 > gpurun_out/ncu_r2c_c.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_vss_env_step --launch-skip 604 -o gpurun_out/prof_r2c_vss1m_serial python tools/step_timing.py --task vss --envs 1048576 --worlds 2 --warmup 300 --no-graph --steps 8 > gpurun_out/ncu_r2c_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
