# oracle shim (TEST INFRASTRUCTURE ONLY): minimal stand-in for `timm`, which the
# reference imports at cra5/models/vaeformer/vit_nlc.py:23 but which is absent here.
