"""CPU oracle for the PC-sampling hot path of szhan227/text2protein.

TEST INFRASTRUCTURE ONLY.  Nothing under ``text2protein_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do,
and there only as the checker (or as the CPU baseline being timed), never as the product path.

What it is: a plain torch-CPU (fp32 weights, fp64 where the reference promotes) restatement of
  * the score UNet forward          (score_sde_pytorch/models/ncsnpp.py:220-263 and the layers it calls)
  * the VE/VP SDE discretisations   (score_sde_pytorch/sde_lib.py:106-157,199-245)
  * the predictor-corrector sampler (score_sde_pytorch/sampling.py:157-289)
plus a numpy Philox4x32-10 / Box-Muller generator matching the in-kernel noise.

Parity pin: the reference ships no tests or golden vectors, so the pins are outputs of
the reference itself, imported unmodified from /root/reference in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks this restatement against every one of them.  Philox is pinned by the Random123 known-answer vectors.
"""
