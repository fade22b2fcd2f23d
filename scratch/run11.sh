for w in 6 10 14; do echo "FL_WINDOW=$w"; FL_WINDOW=$w timeout 200 python profiles/multiseq_bench.py 8 16 160 2>&1 | grep "n_seqs [18]:"; done
