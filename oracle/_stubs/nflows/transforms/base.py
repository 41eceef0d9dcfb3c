class InputOutsideDomain(Exception):
    """Exception to be thrown when the input to a transform is not within its domain."""
