# round 2, session x: compute-sanitizer on the final build (host-step paths included); GPU suite twice (flakiness)
exec > gpurun_out/session_r2x.log 2>&1
set -x
timeout 300 python tools/sanitize_run.py | tail -1
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -4
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | tail -4
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
