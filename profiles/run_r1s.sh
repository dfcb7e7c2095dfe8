# compute-sanitizer over one small pass of every kernel (tools/sanitize_smoke.py), bounded
set -x
timeout 60 python tools/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1; echo rc=$?; tail -2 gpurun_out/sanitize_plain.log
timeout 120 compute-sanitizer --tool memcheck --kernel-name kns=_ZN2yb --print-limit 20 --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/sanitize_memcheck.log
timeout 150 compute-sanitizer --tool racecheck --kernel-name kns=_ZN2yb --print-limit 20 --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/sanitize_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/sanitize_racecheck.log
