import pytest

pytestmark = pytest.mark.gpu


def test_smoke_entry(ge):
    ge.smoke()
