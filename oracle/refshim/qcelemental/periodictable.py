_SYMBOLS = [
    "X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne",
    "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar",
]
# Most abundant isotope masses (NIST 2013 isotopic compositions), as qcelemental>=0.25.1.
_MASS = {
    1: 1.00782503223, 2: 4.00260325413, 3: 7.0160034366, 4: 9.012183065,
    5: 11.00930536, 6: 12.0, 7: 14.00307400443, 8: 15.99491461957,
    9: 18.99840316273, 10: 19.9924401762, 11: 22.9897692820, 12: 23.985041697,
    13: 26.98153853, 14: 27.97692653465, 15: 30.97376199842, 16: 31.9720711744,
    17: 34.968852682, 18: 39.9623831237,
}


def to_mass(z):
    return _MASS[int(z)]


def to_symbol(z):
    return _SYMBOLS[int(z)]


def to_atomic_number(symbol):
    return _SYMBOLS.index(str(symbol).capitalize())
