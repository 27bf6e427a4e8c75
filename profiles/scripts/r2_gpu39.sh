set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
AVIFGPU_MEASURE_ONLY="encode Gray8" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
echo done
