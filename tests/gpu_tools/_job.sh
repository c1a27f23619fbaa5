S=$(date +%s)
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_tools/gpu_sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$? t=$(( $(date +%s)-S ))s"; tail -4 gpurun_out/sanitize_memcheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python tests/gpu_tools/gpu_sanitize.py > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$? t=$(( $(date +%s)-S ))s"; tail -4 gpurun_out/sanitize_synccheck.log
