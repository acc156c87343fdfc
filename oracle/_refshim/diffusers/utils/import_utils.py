def is_xformers_available() -> bool:
    return False
