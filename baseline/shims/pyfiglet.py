def figlet_format(text, font=None):
    return text
