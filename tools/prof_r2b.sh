ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-keccak --no-compiled-cfg3 > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_groestl_leaves" -s 1 -c 2 -o /tmp/gl python tools/mk_prof.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/gl.ncu-rep gpurun_out/r2_merkle_k_groestl_leaves_ncu_full.csv
