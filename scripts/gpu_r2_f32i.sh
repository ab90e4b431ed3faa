#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_nets.py -q -k "fp32_tier_gated" 2>&1 | grep -E "assert 0\.|passed|failed|FAILED" | head
