timeout 300 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -4
echo "== lean"; timeout 200 python scripts/step_sweep_probe.py 2>&1 | grep "'visits': True, 'mode': 'dense_f32'"
echo "== general"; COLO_STEP_KERNEL=general timeout 200 python scripts/step_sweep_probe.py 2>&1 | grep "'visits': True, 'mode': 'dense_f32'"
