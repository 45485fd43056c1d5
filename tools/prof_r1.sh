# Round-1 profiling pass (run on the GPU box via gpurun; outputs land in gpurun_out/).
#  1. launch list of the default bench command (gpu__time_duration per launch)
#  2. ncu --set full of the dominant kernels
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lerp_tma -s 4 -c 2 -o gpurun_out/fold_tma python bench.py --steps 3 --warmup 3 --no-cpu --no-ntt > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_bs -s 4 -c 2 -o gpurun_out/ntt_s1 python tools/ntt_prof.py S1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand_k64 -s 8 -c 4 -o gpurun_out/expand python tools/te_prof.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_tc -s 4 -c 2 -o gpurun_out/roundevals_tc python tools/re_prof.py > /dev/null 2>&1
ls -la gpurun_out
