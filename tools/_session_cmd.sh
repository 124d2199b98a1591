# what the driver runs at round end, on the committed tree: smoke(), then the default bench line
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py | tail -1 | cut -c1-400
