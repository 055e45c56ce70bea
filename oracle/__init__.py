"""CPU oracle for the URSABench SG-MCMC / SWAG / BMA hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``ursabench_b200`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
(or as the CPU baseline being timed), never as the thing shipped.

Layout
------
``restate.py``    numpy restatement of every reference function on the path
                  (each function cites the reference file:line it follows).
``port_torch.py`` torch-CPU port issuing the same op sequence as the reference
                  (per-tensor loops, ``randn_like``): the timed CPU baseline.
``stubs.py`` / ``ref_import.py``  import the live reference from
                  ``/root/reference`` (exists only in the build container).
``gen_golden.py`` runs the live reference with fixed seeds / injected noise and
                  writes the fixtures committed under ``tests/golden/``.

Parity status: PINNED for SGLD/SGHMC/cSGHMC updates, SWA moments, the
covariance ring, smoothing/entropy, ``Prediction`` accumulation and metrics
(golden vectors generated from the unmodified reference, plus the notebook
learning-rate known-answer trace).  The SWAG *draw* has no runnable reference
(reference bug) -> its restatement is the spec, pinned only on
the pieces the reference does execute (``torch.normal`` diag draw formula).
HMC: parity unpinned (third-party hamiltorch, absent).
"""
