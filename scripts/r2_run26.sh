#!/bin/bash
timeout 600 python bench.py --model mmformer --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mmformer 80^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])"
timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mmformer 128^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])"
timeout 300 python eval.py --model rfnet --synthetic 1 --savepath /tmp/eval 2>&1 | tail -3 | cut -c1-200
