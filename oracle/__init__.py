"""CPU oracle for the two-stage pseudo-healthy synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``healthivert-gan_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the CPU baseline.

Each module restates one part of the reference (file:line cited per function) with
torch-CPU fp32 ops for the floating-point rows and numpy for the integer rows.

Pinning status: the reference ships no tests / golden vectors, so the
restatement is pinned against *outputs of the unmodified reference modules imported in
the build container* (``oracle/refshim.py`` + ``oracle/make_golden.py`` ->
``tests/golden/*.npz``) and against the RHLV known answers on the reference's shipped
label volumes.
"""
