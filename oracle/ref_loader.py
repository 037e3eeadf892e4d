"""Import the real, unmodified reference (build container only).

Test infrastructure (see oracle/__init__.py).  ``/root/reference`` is read-only
and does not exist on the GPU box; there the unmodified copy under ``baseline/_ref``
(oracle/install_reference.py) is used.  This module serves (a) ``oracle/make_golden.py``
to generate ``tests/golden/*.npz``, (b) the CPU tests that pin the restatement against
the reference and (c) ``bench.py``'s reference arm / cpu_baseline leg.  The GUI / gym packages the reference imports at module level are
replaced by the empty stand-ins under ``oracle/_stubs``;
they never execute on the headless path.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# the build container has the reference checkout; the GPU box only the copy oracle/install_reference.py placed under
# baseline/_ref (git-ignored, travels with the snapshot)
_CANDIDATES = [os.environ.get('COBEL_REFERENCE_SRC'), '/root/reference/src', os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')]
REFERENCE_SRC = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, 'cobel'))), '/root/reference/src')
_STUBS = os.path.join(_HERE, '_stubs')


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'cobel'))


def load():
    """Return the imported reference package ``cobel`` (raises if unavailable)."""
    if not available():
        raise ImportError('reference sources not found at %s' % REFERENCE_SRC)
    for p in (REFERENCE_SRC, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import cobel  # noqa: F401
    import cobel.agent, cobel.interface, cobel.memory, cobel.policy  # noqa: F401,E401
    import cobel.memory.utils.metrics  # noqa: F401
    import cobel.misc.gridworld_tools, cobel.misc.topology_tools  # noqa: F401,E401
    return cobel
