"""Test infrastructure: CPU restatement of the reference's EM path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py`` may import anything from this package (the product in
``stellarscope_b200/`` never does).
"""
