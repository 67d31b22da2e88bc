from dhd_b200.compat import NECKS, BaseModule


@NECKS.register_module(force=True)
class Identity(BaseModule):
    """Pass-through neck (reference: models/necks/identity.py)."""

    def forward(self, x):
        return x
