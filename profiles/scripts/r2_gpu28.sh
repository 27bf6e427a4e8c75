set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
AVIFGPU_MEASURE_ONLY="premultiplied" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
AVIFGPU_MEASURE_ONLY="straight" python profiles/measure_generic_paths.py 2>/dev/null | cut -c1-200
echo done
