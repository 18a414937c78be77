#!/bin/bash
# SMs the dense persistent kernel leaves free in the chunked host pipeline (CB200_SM_RESERVE; tools/e2e_probe.py)
for r in ${@:-0 1 2 3 4 8}; do echo "CB200_SM_RESERVE=$r"; CB200_SM_RESERVE=$r python tools/e2e_probe.py 2>&1 | sed -n '1p;4p'; done
