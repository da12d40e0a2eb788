timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_c.log | tail -1 | cut -c1-120
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
