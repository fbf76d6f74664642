"""freqair-b200: B200-native (sm_100a) kernels behind the restoration-network forward/backward hot path of
stcodeer/Frequency-wised_All-in-One_Image_Restoration_Model, exposed through the reference's own ``net/``
module API.  See DESIGN.md / INTEGRATION.md at the repository root."""
__version__ = '0.1'
