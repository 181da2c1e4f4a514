"""Host-side drop-in surface that needs no GPU: the reference's import name and the key forms fit() accepts."""
import numpy as np
import pytest


def test_reference_import_name_and_key_forms():
    """`import ComputationalHypergraphDiscovery as CHD` (reference __init__.py:26-29) resolves to this package, and
    fit(key=...) takes the key forms the reference takes (MAIN:275): uint32[2] array-likes, typed-key objects."""
    import ComputationalHypergraphDiscovery as REF
    import ComputationalHypergraphDiscovery.Modes as RM
    from ComputationalHypergraphDiscovery.decision import MinNoiseKernelChooser
    import computationalhypergraphdiscovery_b200 as CHD
    assert REF.GraphDiscovery is CHD.GraphDiscovery and RM.LinearMode is CHD.Modes.LinearMode
    assert MinNoiseKernelChooser is CHD.decision.MinNoiseKernelChooser
    as_key = CHD.random.as_key
    k = CHD.random.PRNGKey(7)
    assert (as_key(None) == CHD.random.PRNGKey(0)).all() and (as_key(7) == k).all()
    assert (as_key([0, 7]) == k).all() and (as_key(np.array([0, 7], dtype=np.int64)) == k).all()

    class TypedKey:                       # quacks like jax.random.key(7): no __array__, raw data behind an attribute
        _base_array = np.array([0, 7], dtype=np.uint32)

        def __array__(self, *a, **kw):
            raise TypeError("typed keys cannot be converted")

    assert (as_key(TypedKey()) == k).all()
    with pytest.raises(TypeError):
        as_key(np.zeros(3))
