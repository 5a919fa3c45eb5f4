#!/bin/bash
for v in 5 4; do echo "== K1 variant $v"; SPRB200_K1_VARIANT=$v python scripts/perf_probe.py 2>&1 | tail -8; SPRB200_K1_VARIANT=$v python scripts/c2_breakdown.py 2>&1 | tail -1; done
