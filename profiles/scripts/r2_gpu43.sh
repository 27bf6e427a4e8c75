# memcheck over the rewritten float decode kernel (the decode-side parity cases and the tuned-kernel shape tests)
timeout 32 compute-sanitizer --tool memcheck --kernel-regex kns=DecodeYccToRgbF32 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py -m gpu -q -x -k "dec_ycc32 or ycc_to_rgb32 or hlg_ootf_exponents" 2>&1 | tail -4
echo memcheck rc=$?
