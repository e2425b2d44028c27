class Calculator:  # placeholder base class for reference interfaces/ase.py:23
    implemented_properties = []

    def __init__(self, *args, **kwargs):
        pass


all_changes = []
