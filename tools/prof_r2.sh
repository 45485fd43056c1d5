# Round-2 profiling pass (run on the GPU box via gpurun; outputs land in gpurun_out/).
#  1. launch lists (gpu__time_duration per launch) of the default bench command and of the compiled keccak replay
#  2. ncu --set full of the kernels that changed this round
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-keccak > gpurun_out/b_under_ncu.log 2>&1
REPLAY_PASSES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_replay.csv ./tools/keccak_replay_cpp 14 > gpurun_out/replay_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lerp_tma -s 4 -c 2 -o gpurun_out/r2_fold_tma python bench.py --steps 3 --warmup 3 --no-cpu --no-ntt --no-keccak > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_tc$ -s 2 -c 2 -o gpurun_out/r2_pair_tc python tools/re_prof.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_uni_b8 -s 2 -c 2 -o gpurun_out/r2_uni_b8 python tools/univariate_bench.py 22 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_groestl -s 2 -c 2 -o gpurun_out/r2_groestl python tools/mk_prof.py > /dev/null 2>&1
ls -la gpurun_out
