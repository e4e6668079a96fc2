#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/determinism_check.py 128 2>&1 | grep -v Warn | sed -n 1,8p
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
