"""Import the real, unmodified reference (build container only).

Test infrastructure (see oracle/__init__.py).  ``/root/reference`` is read-only
and does not exist on the GPU box, so this module is used only (a) by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and (b) by CPU tests
that pin the restatement against the reference (skipped when the reference is
absent).  The GUI / gym packages the reference imports at module level are
replaced by the empty stand-ins under ``oracle/_stubs``;
they never execute on the headless path.
"""
import os
import sys

REFERENCE_SRC = os.environ.get('COBEL_REFERENCE_SRC', '/root/reference/src')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_stubs')


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
