"""TEST SUPPORT: placeholder for the `ray` package.  The reference imports ray at module load
(ramba/ramba_queue.py:17, ramba/ramba.py:135-138) even in RAMBA_NON_DIST mode, which never calls it
(SURVEY.md §8c, Appendix C).  Only oracle/ref_runner.py and tests/golden/make_golden.py put this
directory on sys.path."""


def get(*a, **k):
    raise RuntimeError("ray stub")


def put(*a, **k):
    raise RuntimeError("ray stub")


def wait(*a, **k):
    raise RuntimeError("ray stub")


def is_initialized():
    return False


def remote(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(f):
        return f

    return deco
