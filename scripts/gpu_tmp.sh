# scratch: the command of the last one-off GPU check of the round (final build: smoke + end-to-end render tests)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 60 python -m pytest tests/test_gpu_render.py -q -x -p no:cacheprovider 2>&1 | tail -2
