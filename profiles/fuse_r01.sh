timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bloch_wide" -s 2>&1 | grep -E "fused_launches|passed|failed|Error|differ" | head -10
