#!/bin/bash
cd "$(dirname "$0")/.."
PREC=split timeout 300 python tools/determinism_check.py 128 2>&1 | grep -v Warn | sed -n 1,8p
PREC=split timeout 300 python tools/determinism_check.py 64 2>&1 | grep -v Warn | sed -n 1,4p
