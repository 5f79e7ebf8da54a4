timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_queries or bounds or thresholds or crown" 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras > gpurun_out/i_1024.json 2> gpurun_out/i_1024.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --queries 128 > gpurun_out/i_128.json 2> gpurun_out/i_128.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --queries 512 > gpurun_out/i_512.json 2> gpurun_out/i_512.err
NNSDP_NO_DMMA_IBP=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --queries 128 > gpurun_out/i_128_simt.json 2> gpurun_out/i_128_simt.err
