O=gpurun_out; mkdir -p $O
python -m pytest tests/test_ehvi3d.py tests/test_ehvi.py -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_e3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q -x > $O/pytest_r1p.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_r1p.log; tail -5 $O/pytest_r1p.log
