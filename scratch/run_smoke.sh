timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
bash scratch/run_cfgs.sh 2>&1 | grep "value"
