#!/bin/bash
for s in 0 148; do echo "PM_STAGGER_SMS=$s"; for l in tinyCapture smallCapture bloxCapture; do PM_STAGGER_SMS=$s python tools/time_rules_only.py $l 2>&1 | grep "with obs"; done; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
