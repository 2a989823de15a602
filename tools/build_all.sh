#!/bin/bash
# Build the product library and the -DCP360_ARRIVE_PER_WARP comparison variant (tools/racecheck_probe.py) in-tree,
# and stage the reference into oracle/_ref when /root/reference is present. Run before every gpurun call.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_perwarp.so CP360_NVCC_EXTRA="-DCP360_ARRIVE_PER_WARP" \
  python -c "import cp360_b200; print(cp360_b200.build_library(force=True))"
