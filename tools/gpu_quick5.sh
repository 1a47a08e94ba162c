#!/bin/bash
for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "us/launch"; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
