"""oracle/vii_oracle.py (numpy restatement of geotiepoints/viiinterpolator.py) against the golden vectors of the
reference's own tests (tests/vii_golden.py)."""
import vii_golden
import vii_oracle


def test_vii_oracle_reproduces_the_reference_test_vectors():
    vii_golden.check_reference_cases(vii_oracle.tie_points_interpolation, vii_oracle.tie_points_geo_interpolation)
