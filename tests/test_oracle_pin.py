"""CPU, build container only: the oracle against the unmodified reference in /root/reference."""
import pytest

from oracle import _reference_import


@pytest.mark.skipif(not _reference_import.available(), reason="/root/reference not on this machine")
def test_oracle_pinned_against_reference():
    from oracle import pin_against_reference

    assert pin_against_reference.check_all(verbose=False) == []
