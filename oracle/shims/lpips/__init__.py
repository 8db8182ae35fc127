"""Stub package named ``lpips`` - TEST INFRASTRUCTURE.  The reference imports it at
module scope (src/attr_functions.py:3); LPIPS itself needs VGG weights that are not
available offline, so constructing it raises."""


class LPIPS:
    def __init__(self, *a, **k):
        raise NotImplementedError("LPIPS needs VGG weights; unavailable offline")
