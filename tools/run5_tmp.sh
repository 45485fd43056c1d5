set -x
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_univariate.py -x -q -k "prepare" > gpurun_out/s4_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s4_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_univariate.py -x -q -k "prepare_finish_equals" > gpurun_out/s4_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s4_racecheck.log
ncu --set full --clock-control none --import-source on -k regex:k_uni_b8 -s 3 -c 2 -o /tmp/prep python tools/univariate_bench.py 22 153 75 split > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prep.ncu-rep gpurun_out/r2_univariate_k_uni_b8_prep_ncu_full.csv
python bench.py > gpurun_out/s4_bench2.json 2> gpurun_out/s4_bench2.err
tail -4 gpurun_out/s4_memcheck.log; tail -4 gpurun_out/s4_racecheck.log
