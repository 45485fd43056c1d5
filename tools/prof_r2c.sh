# Round-2 profiling pass of the two-halves univariate round (run on the GPU box via gpurun).
set -x
python tools/univariate_bench.py 24 153 75 split > gpurun_out/s4_uni_split.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
	local name=$1 rx=$2 skip=$3 cnt=$4
	shift 4
	ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o /tmp/$name "$@" > /dev/null 2>&1
	python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/r2_${name}_ncu_full.csv
}
cap univariate_k_uni_finish k_uni_finish 1 1 python tools/univariate_bench.py 22 153 75 split
cap univariate_k_uni_b8_prep k_uni_b8 3 2 python tools/univariate_bench.py 22 153 75 split
for lc in 3 4 5; do REPLAY_PASSES=2 REPLAY_LOG_CHUNKS=$lc ./tools/keccak_replay_cpp 18 > gpurun_out/s4_replay_lc$lc.json 2>&1; done
REPLAY_PASSES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_replay_split.csv ./tools/keccak_replay_cpp 14 > /dev/null 2>&1
cat gpurun_out/s4_uni_split.log
