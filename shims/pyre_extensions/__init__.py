def none_throws(optional, message: str = "Unexpected `None`"):
    if optional is None:
        raise AssertionError(message)
    return optional
