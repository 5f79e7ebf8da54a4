# same-box A/B of the FP64 tensor-core paths (alternate libraries through NNSDP_B200_LIB, switches through the environment)
timeout 900 python -m pytest tests -m gpu -x -q -k "many_queries or bounds or crown or affine or thresholds" 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-extras > gpurun_out/dg_$name.json 2> gpurun_out/dg_$name.err
  env "$@" timeout 300 python tools/crown_timing.py 2>&1 | grep W1000 > gpurun_out/dg_$name.crown
}
run new X=1
run noibp NNSDP_NO_DMMA_IBP=1
run dk32 NNSDP_B200_LIB=/root/repo/nn-sdp_b200/lib/alt_dgemm_dk32.so
