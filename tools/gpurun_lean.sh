#!/bin/bash
# gpurun with a lean snapshot (kernel experiments): the staged demo data / HM binaries / golden vectors stay behind.
# usage: tools/gpurun_lean.sh <timeout_s> '<command>'
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/gpurunignore.keep
trap 'cp /tmp/gpurunignore.keep .gpurunignore' EXIT
printf 'oracle/_ref/data/\noracle/_ref/hm/\ntests/golden/\nprofiles/\n' >> .gpurunignore
/usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
