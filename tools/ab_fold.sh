# A/B of the fold engines on one box: Karatsuba-64 (default) vs 16 x LDS.128
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for e in tma k64 lut128; do
  echo "== engine $e"
  B200_FOLD_ENGINE=$e python bench.py --no-ntt --no-cpu --steps 20 --warmup 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
done
