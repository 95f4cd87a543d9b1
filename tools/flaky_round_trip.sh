for mode in 1 0; do
  fails=0
  for i in $(seq 1 8); do
    VMGB_PDL=$mode timeout 120 python -m pytest tests/test_model_gpu.py -q -m gpu -k round_trip -x 2>&1 | grep -q "1 passed" || fails=$((fails+1))
  done
  echo "PDL=$mode fails=$fails/8"
done
