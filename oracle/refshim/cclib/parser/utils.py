def convertor(value, fromunits, tounits):  # pragma: no cover - never on the hot path
    if fromunits == tounits:
        return value
    raise NotImplementedError("cclib stub: unit conversion is not available")
