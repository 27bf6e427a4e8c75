set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
AVIFGPU_MEASURE_ONLY="32f (a1" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
echo done
