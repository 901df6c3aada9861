#!/bin/bash
# usage: tools/probe.sh "<env assignments>" <bench args...>   -> one line: G ch-samples/s, ms/step, per-step kernel ms
envs="$1"; shift
env $envs python bench.py "$@" --no-cpu-baseline --no-e2e --no-scan-side --no-block-calls --sustain-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-40s %8.2f G  %.4f ms  %s' % ('$envs', d['value']/1e9, d['ms_per_step'], [(k['kind'],round(k['avg_ms'],4)) for k in d['roofline']['kernels']]))"
