O=gpurun_out
python tools/sanitize_small.py > $O/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > $O/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds" $O/sanitize_memcheck.log; tail -4 $O/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/sanitize_small.py > $O/sanitize_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 $O/sanitize_initcheck.log
