mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_initcheck.log 2>&1; echo initcheck rc=$?; tail -4 gpurun_out/sanitize_initcheck.log
