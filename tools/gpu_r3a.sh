#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 120 python tools/time_cbam_pool.py 2>&1 | tail -8
