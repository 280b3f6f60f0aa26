"""Code shared by the two drop-in variants (``epsilonparam/modules`` and ``xparam/modules``)."""
