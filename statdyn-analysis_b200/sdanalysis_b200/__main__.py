from .main import sdanalysis

sdanalysis()
