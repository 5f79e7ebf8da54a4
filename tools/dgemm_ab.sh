# same-box A/B of the FP64 tensor-core paths (alternate libraries through NNSDP_B200_LIB, switches through the environment)
timeout 900 python -m pytest tests -m gpu -x -q -k "gram or many_queries or packed or crown" 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 2 --warmup 3 --no-e2e --no-extras --crown-queries 16 > gpurun_out/gr_$name.json 2> gpurun_out/gr_$name.err
}
run new X=1
run old NNSDP_B200_LIB=/root/repo/nn-sdp_b200/lib/alt_gram_old.so
