# r2ao: reverb pre-delay with one slot register and per-channel reads (fewer live registers in the Cathedral shape)
timeout 300 python profiles/tools/time_reverb_predelay.py 2>&1 | tail -6
timeout 600 python -m pytest tests -m gpu -x -q -k "reverb or Cathedral or preset or golden or full_size or copies" 2>&1 | tail -2
