run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-extras > gpurun_out/t_$name.json 2> gpurun_out/t_$name.err
  env "$@" timeout 300 python tools/crown_timing.py 2>&1 | grep -E "W1000|W100-" > gpurun_out/t_$name.crown
}
run base X=1
run tn64 NNSDP_B200_LIB=/root/repo/nn-sdp_b200/lib/alt_dgemm_tn64.so
NNSDP_B200_LIB=/root/repo/nn-sdp_b200/lib/alt_dgemm_tn64.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "crown or many_queries or lambda" 2>&1 | tail -2
