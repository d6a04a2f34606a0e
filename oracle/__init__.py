"""CPU oracle for the speculative-decoding draft-and-verify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: the only
allowed importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The product package
(``llmspeculativesampling_b200``) never imports this package and has no CPU fallback.

Contents
--------
``ref_ops``       restatement (torch-on-CPU, one row at a time, same ATen op chain) of
                  ``/root/reference/sampling/utils.py:152-245`` and of the accept / resample
                  step of ``/root/reference/sampling/speculative_sampling.py:1966-2023``,
                  with ``torch.multinomial`` replaced by an explicit inverse-CDF rule driven
                  by pre-drawn uniforms (BASELINE.json north_star parity contract).
``tape``          the per-request uniform tape (2*gamma+2 uniforms per iteration).
``replay_model``  a fake causal LM whose logits depend only on the token prefix, so the
                  reference (batch 1, KV tuple cache) and the B200 engine (batched, static KV)
                  see bit-identical logits.
``spec_loop``     restatement of the whole ``speculative_sampling`` loop
                  (``speculative_sampling.py:1934-2043``) on top of ``ref_ops``.
``ref_loader``    imports the UNMODIFIED reference from ``/root/reference`` (dev container only;
                  that path does not exist on the GPU box) through two compatibility shims and
                  patches its RNG with the tape.  Used by ``make_golden.py`` and by the
                  ``-m "not gpu"`` tests that pin the restatement against the real reference.

Parity pin status: the reference ships NO tests, golden vectors or fixtures, so
parity is *unpinned by the reference's own tests*; the pins used here are outputs of the
reference code itself, run in the dev container by ``oracle/make_golden.py`` and committed
under ``tests/golden/``.
"""
