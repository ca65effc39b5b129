for v in ${VARIANTS:-0 1 2 3}; do
  echo "=== variant $v"
  TISEG_FLOOD_VARIANT=$v python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "watershed_u8 or postproc_dist" 2>&1 | grep -E "passed|failed|Error|assert|mismatch" | head -5
  TISEG_FLOOD_VARIANT=$v python scripts/flood_debug.py 2>&1 | grep -E "tiseg flood|k_ws_flood" | tail -4
done
