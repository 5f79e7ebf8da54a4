set -x
timeout 300 python tools/sanitize_smoke.py > gpurun_out/san_plain.log 2>&1; tail -2 gpurun_out/san_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/san_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/san_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/san_racecheck.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo bench rc=$?
