/* b200_mapper.h — C entry point of the B200 placement policy for Regent-FFT (see b200_mapper.cc).
 * Same shape as the reference's test/test_mapper.h:23 so that a Regent program registers it the same way:
 *     local cmapper = terralib.includec("b200_mapper.h")  ...  regentlib.start(main, cmapper.register_mappers)
 */
#ifndef FFT_B200_MAPPER_H
#define FFT_B200_MAPPER_H

#ifdef __cplusplus
extern "C" {
#endif

void register_mappers(void);

#ifdef __cplusplus
}
#endif

#endif /* FFT_B200_MAPPER_H */
