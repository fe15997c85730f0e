"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package ``nesvor_b200``).

CPU restatement of the NeSVoR reconstruction hot path, used as the parity checker for the CUDA
kernels and as the CPU baseline in ``bench.py``:

* ``nsv_oracle.c`` (+ ``*_impl.h``)  plain-C restatement of ``slice_acquisition`` (forward, backward,
  adjoint, adjoint-backward, equalize) and of ``axisangle2mat`` / ``mat2axisangle`` (fwd + bwd).
  Pinned bit-for-bit against ``oracle/_ref`` (the reference's own kernel bodies compiled for CPU by
  ``build_ref.sh``), against the reference's 11 scipy golden vectors and against its CG
  known-answer test.
* ``inr_oracle.py``  pure-PyTorch fp32/fp64 restatement of the INR path (hash-grid encoding, MLPs,
  ``NeSVoR.forward`` losses, AdamW step).  The hash grid and fp16 MLP of the reference live in
  tiny-cuda-nn (unpinned pip-from-git dependency, absent from /root/reference and not installable
  here), so **parity for hash-encode+MLP is unpinned**: the oracle restates tcnn's published
  algorithm (Instant-NGP, Mueller et al. 2022, eqs. 2-4 + SURVEY.md App. A) and is anchored on the
  reference's own call sites (nesvor/nesvor/models.py:22-41,142-152).  Everything around it
  (models.py:260-384 losses / regularisers, transform.py point transforms) is restated op-for-op.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.
"""
