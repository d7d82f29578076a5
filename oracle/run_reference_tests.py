"""
Run the reference's own emulator tests against the UNMODIFIED reference imported through
oracle.refshim (build container only; /root/reference is absent on the GPU box).

    python -m oracle.run_reference_tests            # reference backends (pins the shim
                                                    #   and the restated thewalrus.perm)
    python -m oracle.run_reference_tests --b200     # same tests with Backend._backend_map
                                                    #   redirected to lightworks_b200 (needs a GPU)

Test infrastructure only.
"""
from __future__ import annotations

import os
import sys
import tempfile


class _ShimPlugin:
    def __init__(self, b200: bool):
        self.b200 = b200

    def pytest_configure(self, config):
        config.addinivalue_line("markers", "flaky: rerun marker (pytest-rerunfailures absent)")
        from oracle import refshim

        refshim.install()
        if self.b200:
            import lightworks_b200

            lightworks_b200.install()


def main(argv=None) -> int:
    import pytest

    argv = list(sys.argv[1:] if argv is None else argv)
    b200 = "--b200" in argv
    argv = [a for a in argv if a != "--b200"]
    from oracle import refshim

    tests_root = os.path.join(refshim.REFERENCE_ROOT, "tests")
    # arguments naming a directory/file under the reference's tests/ (e.g. "qubit/gate_test.py") become targets,
    # everything else is passed to pytest unchanged
    targets, rest = [], []
    for a in argv:
        cand = a if os.path.isabs(a) else os.path.join(tests_root, a)
        (targets if (not a.startswith("-") and os.path.exists(cand)) else rest).append(cand if os.path.exists(cand) and not a.startswith("-") else a)
    targets = targets or [os.path.join(tests_root, "emulator")]
    with tempfile.TemporaryDirectory() as tmp:
        return pytest.main(
            [*targets, "-q", "-p", "no:cacheprovider", f"--rootdir={tmp}", "-c", os.devnull,
             "-W", "ignore", *rest],
            plugins=[_ShimPlugin(b200)],
        )


if __name__ == "__main__":
    sys.exit(main())
