timeout 300 python profiles/multiseq_phase_times.py 8 8 160 2>&1 | tail -28
