#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/determinism_check.py 128 2>&1 | grep -v Warn | tail -24
