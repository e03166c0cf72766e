"""Empty stand-in: the reference imports matplotlib.pyplot unused (lib/algorithms/advanced/simple_zeroshot_opt.py:3)."""
