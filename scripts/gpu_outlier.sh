#!/bin/bash
for v in mma tc3; do echo "=== DRAG_ATTENTION=$v"; DRAG_ATTENTION=$v timeout 300 python scripts/outlier_debug.py outlier all 2>&1 | tail -n 9; done
