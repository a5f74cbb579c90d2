#!/bin/bash
# A/B of two builds of the library on the same box: build/libA.so (current) vs build/libB.so (weight gradient of commit da7ede6)
cp passion_b200/libpassion_b200.so build/libCur.so
for v in B A B A; do
cp build/lib$v.so passion_b200/libpassion_b200.so
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['families_ms_per_step']['conv3d_wgrad_tc'])"
done
cp build/libCur.so passion_b200/libpassion_b200.so
