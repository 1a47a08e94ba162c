#!/bin/bash
for e in 32 8; do
  echo "== PM_ENVS_PER_CTA=$e"
  for n in 1024 4096 8192 16384; do PM_ENVS_PER_CTA=$e python tools/time_rules_only.py tinyCapture $n 2>&1 | grep "us/launch" | tr '\n' ' '; echo; done
  PM_ENVS_PER_CTA=$e python tools/time_rules_only.py bloxCapture 2048 2>&1 | grep "with obs"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
