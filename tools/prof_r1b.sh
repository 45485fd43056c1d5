set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lerp_lut -s 4 -c 2 -o gpurun_out/fold python bench.py --steps 3 --warmup 3 --no-cpu --no-ntt > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_bs -s 4 -c 2 -o gpurun_out/ntt_s1 python tools/ntt_prof.py S1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 30 -c 3 -o gpurun_out/expand python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
