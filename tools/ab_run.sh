#!/bin/bash
# On the GPU box: time the bench workload (device-resident arm only) for each kernel variant x envs-per-block.
# usage: tools/ab_run.sh "<variant names>" "<E values>"
cd "$(dirname "$0")/.."
for v in $1; do for e in $2; do
  FM_B200_LIB=$PWD/factory_marl_b200/csrc/build/variants/$v.so timeout 200 python bench.py --no-cpu --no-ppo --no-cfg3 --no-sync-window --steps 10 --warmup 3 --envs-per-block $e 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v E=$e value %.0f e2e %.0f kernel_ms %.2f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms']))"
done; done
