/* empty stand-in so that the reference's beamformer.hh can be included without -DDEBUG
 * (observation-mode constants) where PSRDADA is not installed.  Test infrastructure only. */
