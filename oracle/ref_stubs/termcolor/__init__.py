def colored(text, *args, **kwargs):
    return text
