from .psf import GAUSSIAN_FWHM, SINC_FWHM, resolution2sigma, get_PSF
from .misc import meshgrid, gaussian_blur, MovingAverage
from .loss import ncc_loss
