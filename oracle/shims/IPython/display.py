def clear_output(*a, **k):
    return None
