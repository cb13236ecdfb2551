"""TEST INFRASTRUCTURE ONLY.  CPU restatement (torch fp32) of the reference algorithm for the Video-ViT forward path,
plus (re-exported from the repo-level synth_data.py) the deterministic synthetic weights / inputs both sides of every parity
test are fed.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package; the
product (simple-tad_b200/) never does.  Parity is PINNED: oracle/make_golden.py imports the unmodified reference
modules from /root/reference, runs them on these synthetic inputs and commits their outputs under tests/golden/;
tests/test_oracle.py checks this restatement against those outputs.
"""
