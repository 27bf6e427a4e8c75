set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
AVIFGPU_MEASURE_ONLY="-> RGB" python profiles/measure_generic_paths.py 2>/dev/null | grep -v 32f | cut -c1-200
echo done
