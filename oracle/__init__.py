"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference SupCon objective.

Nothing in the product package (``wav2vec_contr_loss_b200``) imports this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or as the timed CPU baseline -- never as the thing shipped.

Parity pinning: the reference repository ships no tests or golden vectors for
this path.  The oracle is therefore pinned by
(1) executing the reference's own ``loss.py`` in the build container
    (``oracle/ref_loader.py`` + ``oracle/gen_golden.py``) and committing the
    outputs under ``tests/golden/``, and
(2) the RNG-free golden table and analytic known-answer tests of SURVEY.md
    Appendix B (also produced from the reference's ``loss.py``).

Modules
  supcon_oracle.py  closed-form (vectorised, row-blocked) restatement + a
                    per-anchor loop port used as the timed CPU baseline
  ref_loader.py     imports /root/reference/loss.py when it exists (build
                    container only; never on the GPU box)
  gen_golden.py     writes tests/golden/*.npz from the real reference
"""
