"""Unit conversion constants of the electrostatic prefactor (values as in the reference's ``prefactors.py``)."""

SI = 2.3070775523417355e-28  #: Gaussian units -> SI
eV_A = 14.399645478425667  #: Gaussian units -> eV / Angstrom
kcalmol_A = 332.0637132991921  #: Gaussian units -> kcal/mol / Angstrom
kJmol = 1389.3545764438197  #: Gaussian units -> kJ/mol / Angstrom
