"""oracle shim (TEST INFRASTRUCTURE ONLY) for the `dict_recursive_update` package
used at /root/reference/cra5/models/vaeformer/vit_nlc.py:18,1027."""


def recursive_update(default, custom):
    if not isinstance(default, dict) or not isinstance(custom, dict):
        raise TypeError("Params of recursive_update should be dicts")
    for key in custom:
        if isinstance(custom[key], dict) and isinstance(default.get(key), dict):
            default[key] = recursive_update(default[key], custom[key])
        else:
            default[key] = custom[key]
    return default
