#!/bin/bash
# A/B of library builds / env switches on ONE box: each argument is "label:ENV=VAL,ENV=VAL" (YB_LIB selects a build).
cd "$(dirname "$0")/.."
for spec in "$@"; do
  label="${spec%%:*}"; envs="${spec#*:}"
  echo "== $label"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 ${BENCH_ARGS} 2>&1 | tail -4 | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
try: d=json.loads(t[-1])
except Exception: print('bench failed:', *t, sep='\n  '); sys.exit(0)
r=d.get('roofline') or {}
print('img/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', r.get('frac'), 'clk', d.get('clocks'))" )
done
