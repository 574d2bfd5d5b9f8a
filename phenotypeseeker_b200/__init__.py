"""B200-native k-mer association path for PhenotypeSeeker (count -> matrix -> chi2/t -> filter)."""
__version__ = "0.1.0"
