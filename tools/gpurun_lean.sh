#!/bin/bash
# gpurun with a lean snapshot (kernel experiments): the staged demo data / HM binaries / golden vectors stay behind.
# (tools/variants is in the default .gpurunignore: lean runs that need variant libraries pass KEEP_VARIANTS=1)
# usage: tools/gpurun_lean.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/gpurunignore.keep
trap 'cp /tmp/gpurunignore.keep .gpurunignore' EXIT
printf 'oracle/_ref/data\noracle/_ref/hm\ntests/golden\nprofiles\n' >> .gpurunignore
if [ -n "$KEEP_VARIANTS" ]; then grep -v '^tools/variants$' .gpurunignore > /tmp/gpurunignore.tmp; cp /tmp/gpurunignore.tmp .gpurunignore; fi
/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"
