# Round-2 profiling pass (run on the GPU box via gpurun; outputs land in gpurun_out/).
#  1. launch lists (gpu__time_duration per launch) of the default bench command and of the compiled keccak replay
#  2. ncu --set full of the kernels that changed this round, condensed on the box (tools/ncu_summary.py): the reports
#     themselves are ~16 MB each and gpurun brings back at most 64 MiB
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-keccak --no-compiled-cfg3 > gpurun_out/b_under_ncu.log 2>&1
REPLAY_PASSES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_replay.csv ./tools/keccak_replay_cpp 14 > gpurun_out/replay_under_ncu.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
	local name=$1 rx=$2 skip=$3 cnt=$4
	shift 4
	ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
	python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/r2_${name}_ncu_full.csv
}
cap fold_k_lerp_tma k_lerp_tma 4 2 python bench.py --steps 3 --warmup 3 --no-cpu --no-ntt --no-keccak --no-compiled-cfg3
cap round_evals_k_pair_tc 'k_pair_tc$' 2 2 python tools/re_prof.py
cap univariate_k_uni_b8 k_uni_b8 2 2 python tools/univariate_bench.py 22
cap merkle_k_groestl k_groestl_leaves 1 2 python tools/mk_prof.py
# (k_sumcheck_tail_grid cannot be captured by kernel replay: it waits for challenges the host posts AFTER the launch call
#  returns, and ncu replays inside that call -- its evidence is the kernel's own time stamps, REPLAY_TAIL_TRACE=1)
ls -la gpurun_out
