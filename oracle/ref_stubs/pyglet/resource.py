model = None
