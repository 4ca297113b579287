"""CPU oracle for the tokenize/detokenize hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain fp32 PyTorch-CPU restatement of the arithmetic the
reference wrappers (`/root/reference/audiocodecs/{codec,encodec,dac,mimi}.py`)
reach through their third-party backends (transformers 5.5.0
`EncodecModel`/`MimiModel`/`DacModel`, descript-audio-codec 1.0.0 for DAC and
torchaudio 2.11 `functional.resample`).  Every function cites the file:line it
follows (`R/` = /root/reference, `HF/` = site-packages/transformers/models,
`TA/` = site-packages/torchaudio/functional/functional.py).

It is NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product package `audiocodecs_b200` never imports it and has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so
the oracle is pinned against the *live* reference -- the unmodified wrapper
classes imported from /root/reference over transformers' own modules -- by
`oracle/make_golden.py`, whose outputs are committed under `tests/golden/`.
"""
